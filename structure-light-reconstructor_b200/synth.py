"""Host-side (numpy) synthetic scans for tests: a rectified stereo view of a plane + bump scene lit
by the reference's own patterns.  Product-side utility; does not touch oracle/."""
from __future__ import annotations

import numpy as np

# first three: Duke/multifrequency.cpp:3.  The fourth is 56, not the 55 a constant-difference series would give: a
# heterodyne cascade subtracts neighbouring beats, (70-64)-(64-59) = 1 and (64-59)-(59-f4) must not be 1 as well, or the
# last level beats at frequency 0 and every pixel of a row decodes to the same phase (with 56: 1 and 2 -> one period).
FREQ = (70, 64, 59, 56, 52, 50, 47, 45)
PI_GEN = 3.1416            # Duke/multifrequency.h:5


def scene_coords(W, H, proj_w, seed=0, integer_disparity=True):
    """Projector column u[cam, i, x] seen by every camera pixel and the lit mask."""
    rng = np.random.default_rng(seed)
    d0 = 24.0 + 16.0 * rng.random()
    d1 = 8.0 * (rng.random() - 0.5) / H
    amp = 6.0 + 10.0 * rng.random()
    cx, cy = W * (0.35 + 0.3 * rng.random()), H * (0.35 + 0.3 * rng.random())
    sig = 0.18 * W
    xs = np.arange(W, dtype=np.float64)[None, :]
    ys = np.arange(H, dtype=np.float64)[:, None]
    d = d0 + d1 * ys + amp * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (sig * sig))
    if integer_disparity:
        d = np.rint(d)
    g_scale, g_off = 0.92 * proj_w / W, 0.02 * proj_w
    k = np.stack([xs - d, np.broadcast_to(xs, (H, W))])           # right-equivalent column
    u = g_scale * k + g_off + 0.02 * ys[None]
    lit = (u >= 0) & (u < proj_w) & (k >= 0)
    for q in range(3):                                            # per-camera shadow rectangles
        x0, y0 = int(rng.random() * 0.85 * W), int(rng.random() * 0.85 * H)
        w, h = int((0.06 + 0.08 * rng.random()) * W), int((0.06 + 0.08 * rng.random()) * H)
        for cam in range(2):
            lit[cam, y0:y0 + h, x0 + 37 * cam:x0 + 37 * cam + w] = False
    return u, lit


def _finish(img, noise_dn, rng):
    if noise_dn > 0:
        img = img + rng.normal(0.0, noise_dn, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_mf(W, H, proj_w=None, seed=0, integer_disparity=True, noise_dn=0.0, F=3, S=4):
    """uint8 [2, 2+F*S, H, W] multi-frequency stack (layout of Duke/multifrequency.cpp:16-17,30 for F=3, S=4)."""
    proj_w = proj_w or W
    u, lit = scene_coords(W, H, proj_w, seed, integer_disparity)
    rng = np.random.default_rng(seed + 1000)
    out = np.empty((2, 2 + F * S, H, W), np.uint8)
    out[:, 0] = _finish(np.where(lit, 200.0, 30.0), noise_dn, rng)
    out[:, 1] = _finish(np.full(lit.shape, 20.0), noise_dn, rng)
    for f in range(F):
        for s in range(S):
            arg = (PI_GEN * 2 * u * FREQ[f] / proj_w + PI_GEN * 2 * s / S).astype(np.float32)
            v = np.trunc(np.float32(135.0) + np.float32(79.0) * np.cos(arg))
            out[:, 2 + S * f + s] = _finish(np.where(lit, v, 20.0), noise_dn, rng)
    return out


def synth_gray(W, H, scan_w=None, scan_h=None, seed=0, integer_disparity=True, noise_dn=0.0, rows=False):
    """uint8 [2, N, H, W] Gray stack (layout of Duke/graycodes.cpp:63-110); rows=True adds row bits."""
    scan_w = scan_w or W
    scan_h = scan_h or H
    nc = int(np.ceil(np.log(float(scan_w)) / np.log(2.0)))
    nr = int(np.ceil(np.log(float(scan_h)) / np.log(2.0))) if rows else 0
    u, lit = scene_coords(W, H, scan_w, seed, integer_disparity)
    rng = np.random.default_rng(seed + 2000)
    N = 2 + 2 * nc + 2 * nr
    out = np.empty((2, N, H, W), np.uint8)
    out[:, 0] = _finish(np.where(lit, 200.0, 30.0), noise_dn, rng)
    out[:, 1] = _finish(np.full(lit.shape, 20.0), noise_dn, rng)
    col = np.where(lit, u, 0).astype(np.int64)
    gray = col ^ (col >> 1)
    for c in range(nc):
        bit = (gray >> (nc - 1 - c)) & 1
        out[:, 2 + 2 * c] = _finish(np.where(lit & (bit == 1), 200.0, 20.0), noise_dn, rng)
        out[:, 3 + 2 * c] = _finish(np.where(lit & (bit == 0), 200.0, 20.0), noise_dn, rng)
    if rows:
        v = np.broadcast_to((np.arange(H, dtype=np.float64)[:, None] * scan_h / H), (H, W))
        rowc = np.broadcast_to(v.astype(np.int64), lit.shape)
        grow = rowc ^ (rowc >> 1)
        for c in range(nr):
            bit = (grow >> (nr - 1 - c)) & 1
            out[:, 2 + 2 * nc + 2 * c] = _finish(np.where(lit & (bit == 1), 200.0, 20.0), noise_dn, rng)
            out[:, 3 + 2 * nc + 2 * c] = _finish(np.where(lit & (bit == 0), 200.0, 20.0), noise_dn, rng)
    return out


def write_png_opencv_style(path, img):
    """An 8-bit grey PNG as cv::imwrite writes the reference's scan images: Sub filter on every row, zlib level 1 with
    the Z_RLE strategy, the stream cut into 8 KB IDAT chunks (libpng's default buffer)."""
    import struct
    import zlib
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    rows = np.empty((h, w + 1), np.uint8)
    rows[:, 0] = 1
    rows[:, 1] = img[:, 0]
    rows[:, 2:] = img[:, 1:] - img[:, :-1]          # uint8 arithmetic wraps mod 256, as the filter does
    c = zlib.compressobj(1, zlib.DEFLATED, 15, 8, zlib.Z_RLE)
    z = c.compress(rows.tobytes()) + c.flush()

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
                b"".join(chunk(b"IDAT", z[o:o + 8192]) for o in range(0, len(z), 8192)) + chunk(b"IEND", b""))
