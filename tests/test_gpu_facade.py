"""GPU tests of the drop-in class surface: a project directory laid out as the reference expects
(calib/*.txt, scan/{left,right}/<sn>/{L,R}<i>.png) is reconstructed by the Qt-free MFReconstruct / Reconstruct
facades (facade_demo drives them like MainWindow::startreconstruct) and the resulting PointCloudImage is
compared with the oracle applied to the same files."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import slr_b200
from slr_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "structure-light-reconstructor_b200", "facade", "facade_demo")


def write_png(path, img):
    h, w = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 1)) + chunk(b"IEND", b""))


def write_mat(path, m):
    m = np.atleast_2d(np.asarray(m, np.float64))
    with open(path, "w") as f:
        for r in m:
            f.write("\t".join(f"{v:.9g}" for v in r) + "\t\n")


def f32(a):
    return np.asarray(a, np.float64).astype(np.float32)


def make_project(tmp, W, H, sn, stacks, rigid=None):
    cams = {}
    for side, fx, fy, cx, cy, dist, R, t in (
            ("left", 600.0, 601.0, W / 2 + 2.25, H / 2 - 1.5, [-0.11, 0.07, 0.0006, -0.0003, 0.0], np.eye(3), [0, 0, 0]),
            ("right", 598.0, 600.5, W / 2 - 1.75, H / 2 + 1.25, [-0.09, 0.04, -0.0004, 0.0005, 0.0],
             [[0.9998, 0.0, 0.02], [0.0, 1.0, 0.0], [-0.02, 0.0, 0.9998]], [-80.0, 0.5, 1.0])):
        d = os.path.join(tmp, "calib", side)
        os.makedirs(d, exist_ok=True)
        K = [[fx, 0, cx], [0, fy, cy], [0, 0, 1]]
        write_mat(os.path.join(d, "cam_matrix.txt"), K)
        write_mat(os.path.join(d, "cam_distortion.txt"), np.array(dist)[:, None])
        write_mat(os.path.join(d, "cam_rotation_matrix.txt"), R)
        write_mat(os.path.join(d, "cam_trans_vectror.txt"), np.array(t)[:, None])
        write_mat(os.path.join(d, "cam_stereo.txt"), K)
        write_mat(os.path.join(d, "distortion_stereo.txt"), np.array(dist)[:, None])
        cams[side] = slr_b200.Camera(fc=tuple(f32([fx, fy])), cc=tuple(f32([cx, cy])), dist=tuple(f32(dist)),
                                     R=tuple(f32(R).ravel()), t=tuple(f32(t)))
    Rs = [[0.99985, 0.002, 0.0172], [-0.0021, 0.999995, 0.004], [-0.0172, -0.004, 0.99984]]
    write_mat(os.path.join(tmp, "calib", "R_stereo.txt"), Rs)
    write_mat(os.path.join(tmp, "calib", "T_stereo.txt"), np.array([-80.0, 0.4, 1.2])[:, None])
    for name in ("fundamental_stereo.txt", "H1_mat.txt", "H2_mat.txt"):
        write_mat(os.path.join(tmp, "calib", name), np.eye(3))
    for cam, side, pre in ((0, "left", "L"), (1, "right", "R")):
        d = os.path.join(tmp, "scan", side, str(sn))
        os.makedirs(d, exist_ok=True)
        for i in range(stacks.shape[1]):
            write_png(os.path.join(d, f"{pre}{i}.png"), stacks[cam, i])
    if rigid is not None:
        write_mat(os.path.join(tmp, "scan", f"transfer_mat{sn}.txt"), rigid)
    return [cams["left"], cams["right"]]


def run_demo(kind, tmp, sn, scan_w, scan_h, W, H, black, white, color, ply=None, autocontrast=False):
    out = os.path.join(tmp, "out.bin")
    env = dict(os.environ)
    env["DUKE_AUTOCONTRAST"] = "1" if autocontrast else "0"
    if ply:
        env["DUKE_EXPORT_PLY"] = ply      # the caller's export step (MeshCreator, mainwindow.cpp:637-646)
    r = subprocess.run([DEMO, kind, tmp, str(sn), str(scan_w), str(scan_h), str(W), str(H), str(black), str(white),
                        str(int(color)), out], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr + r.stdout
    buf = open(out, "rb").read()
    w, h = struct.unpack_from("<ii", buf, 0)
    o = 8
    sums = np.frombuffer(buf, np.float32, w * h * 3, o).reshape(h, w, 3)
    o += w * h * 12
    cnt = np.frombuffer(buf, np.uint8, w * h, o).reshape(h, w)
    o += w * h
    (has_sr,) = struct.unpack_from("<i", buf, o)
    o += 4
    Q = m1 = m2 = None
    if has_sr:
        Q = np.frombuffer(buf, np.float64, 16, o).reshape(4, 4)
        o += 128
        m1 = np.frombuffer(buf, np.int16, 2 * H * W * 2, o).reshape(2, H, W, 2)
        o += 2 * H * W * 4
        m2 = np.frombuffer(buf, np.uint16, 2 * H * W, o).reshape(2, H, W)
        o += 2 * H * W * 2
    probes = np.frombuffer(buf, np.float32, 64, o).reshape(16, 4)
    return sums, cnt, Q, m1, m2, probes


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


RIGID = [[0.98, -0.17, 0.05, 12.5], [0.17, 0.98, 0.02, -3.25], [-0.05, -0.01, 0.99, 40.0]]


@pytest.mark.parametrize("sn,rigid", [(0, None), (3, RIGID)])
def test_mfreconstruct_facade_end_to_end(tmp_path, oracle, sn, rigid):
    W, H = 320, 64
    scan_w, scan_h = 320, 64        # as in the reference's defaults scan == camera size: F7 drops columns >= scan_h
    stacks = synth.synth_mf(W, H, seed=61, noise_dn=1.0)
    cams = make_project(str(tmp_path), W, H, sn, stacks, rigid)
    ply = str(tmp_path / "cloud.ply")
    sums, cnt, Q, m1, m2, probes = run_demo("mf", str(tmp_path), sn, scan_w, scan_h, W, H, 40, 0, False, ply=ply)
    # startreconstruct's last step: MeshCreator(points3DProjView).exportPlyMesh == the reference's file for this cloud
    oracle.export_mesh(sums, cnt, scan_w, scan_h, tmp_path / "oracle.ply", False, None)
    assert open(ply, "rb").read() == open(tmp_path / "oracle.ply", "rb").read()
    rect = np.stack([[oracle.remap_linear(stacks[c, i], m1[c], m2[c]) for i in range(14)] for c in range(2)])
    xyz, valid, k, n = oracle.run_mf(rect, cams, Q, rigid=f32(rigid) if rigid else None)
    pts_o, cnt_o = oracle.pointcloud_from_dense(xyz, valid, scan_w, scan_h)
    assert n > 1000 and (cnt == cnt_o).all() and cnt.sum() > 100
    assert (bits(sums[cnt > 0]) == bits(pts_o[cnt_o > 0])).all()
    # stereoRect::Q has the cv::stereoRectify layout
    assert Q[0, 0] == 1 and Q[1, 1] == 1 and Q[2, 3] > 100 and Q[3, 2] > 0
    for k_ in range(16):   # getPoint probes: mean = sum * (1.f / count)
        i, j = (k_ * 7919) % scan_w, (k_ * 104729) % scan_h
        if cnt[j, i]:
            assert probes[k_, 0] == 1 and np.allclose(probes[k_, 1:], sums[j, i] / cnt[j, i], rtol=1e-6)
        else:
            assert probes[k_, 0] == 0


def _write_png_filtered(path, img, mode, rng):
    """8-bit grey PNG the way real encoders write it: per-row filters (mode 'sub' = OpenCV's encoder: Sub on every row,
    Z_RLE; 'up' = None / Sub / Up rows; 'all' = all five filter types) and the zlib stream cut into 8 KB IDAT chunks."""
    h, w = img.shape
    a = img.astype(np.int32)
    rows = []
    for y in range(h):
        t = {"sub": 1, "up": int(rng.integers(0, 3)), "all": int(rng.integers(0, 5))}[mode]
        left = np.concatenate([[0], a[y, :-1]])
        up = a[y - 1] if y else np.zeros(w, np.int32)
        ul = np.concatenate([[0], up[:-1]])
        if t == 0:
            f = a[y]
        elif t == 1:
            f = a[y] - left
        elif t == 2:
            f = a[y] - up
        elif t == 3:
            f = a[y] - ((left + up) >> 1)
        else:
            pa, pb, pc = np.abs(up - ul), np.abs(left - ul), np.abs(left + up - 2 * ul)
            f = a[y] - np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
        rows.append(bytes([t]) + (f & 255).astype(np.uint8).tobytes())
    c = zlib.compressobj(1, zlib.DEFLATED, 15, 8, zlib.Z_RLE if mode == "sub" else zlib.Z_DEFAULT_STRATEGY)
    z = c.compress(b"".join(rows)) + c.flush()

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
                chunk(b"tEXt", b"Software\x00test") + b"".join(chunk(b"IDAT", z[o:o + 8192]) for o in range(0, len(z), 8192)) +
                chunk(b"IEND", b""))


def test_mfreconstruct_ingests_real_world_png_encodings(tmp_path, oracle):
    """The ingest path (facade/ingest.cpp + k_png.cu) on the encodings scan images come in: OpenCV-style Sub-filtered
    files go to the GPU as filtered scanlines, files with Up rows too (second kernel), files with Average / Paeth rows
    are unfiltered on their host thread, a .pgm stands in for a missing .png.  Same cloud as the oracle on the pixels,
    and as the host-decode route (DUKE_HOST_DECODE=1: zlib-free reader + slr_run_mf_host + addDense).  scan size !=
    camera size exercises addPoint's (i_w, j_h) drop rule in the GPU epilogue."""
    W, H = 320, 64
    scan_w, scan_h = 48, 400
    stacks = synth.synth_mf(W, H, seed=67, integer_disparity=False, noise_dn=2.0)
    cams = make_project(str(tmp_path), W, H, 0, stacks)
    rng = np.random.default_rng(1)
    for cam, side, pre in ((0, "left", "L"), (1, "right", "R")):
        for i in range(14):
            path = os.path.join(str(tmp_path), "scan", side, "0", f"{pre}{i}.png")
            mode = ("sub", "up", "all")[(i + cam) % 3]
            _write_png_filtered(path, stacks[cam, i], mode, rng)
    os.remove(os.path.join(str(tmp_path), "scan", "right", "0", "R5.png"))
    with open(os.path.join(str(tmp_path), "scan", "right", "0", "R5.pgm"), "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (W, H) + stacks[1, 5].tobytes())
    sums, cnt, Q, m1, m2, _ = run_demo("mf", str(tmp_path), 0, scan_w, scan_h, W, H, 40, 0, False)
    rect = np.stack([[oracle.remap_linear(stacks[c, i], m1[c], m2[c]) for i in range(14)] for c in range(2)])
    xyz, valid, k, n = oracle.run_mf(rect, cams, Q)
    pts_o, cnt_o = oracle.pointcloud_from_dense(xyz, valid, scan_w, scan_h)
    assert cnt.shape == (scan_h, scan_w) and (cnt == cnt_o).all() and cnt.sum() > 500
    assert (bits(sums[cnt > 0]) == bits(pts_o[cnt_o > 0])).all() and (sums[cnt == 0] == 0).all()
    os.environ["DUKE_HOST_DECODE"] = "1"
    try:
        sums_h, cnt_h, *_ = run_demo("mf", str(tmp_path), 0, scan_w, scan_h, W, H, 40, 0, False)
    finally:
        del os.environ["DUKE_HOST_DECODE"]
    assert (cnt_h == cnt).all() and (bits(sums_h) == bits(sums)).all()
    # a damaged file is reported, not decoded
    p = os.path.join(str(tmp_path), "scan", "left", "0", "L3.png")
    blob = bytearray(open(p, "rb").read())
    blob[len(blob) // 2] ^= 0x40
    open(p, "wb").write(bytes(blob))
    out = os.path.join(str(tmp_path), "out2.bin")
    r = subprocess.run([DEMO, "mf", str(tmp_path), "0", str(scan_w), str(scan_h), str(W), str(H), "40", "0", "0", out],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "L3.png" in r.stderr


def test_reconstruct_ge_facade_end_to_end(tmp_path, oracle):
    W, H = 320, 48
    stacks = synth.synth_gray(W, H, seed=62, noise_dn=2.0)
    nc = oracle.gray_num_bits(W)
    make_project(str(tmp_path), W, H, 0, stacks)
    sums, cnt, Q, m1, m2, _ = run_demo("ge", str(tmp_path), 0, W, 400, W, H, 40, 4, True)
    rect = np.stack([[oracle.remap_linear(stacks[c, i], m1[c], m2[c]) for i in range(stacks.shape[1])] for c in range(2)])
    dec = [oracle.gray_decode(rect[c], nc, 0, 40, 4, W, 400) for c in range(2)]
    xyz, valid, k, color, n = oracle.ge_triangulate(dec[0][0], dec[0][2], dec[1][0], dec[1][2], Q,
                                                    whiteL=rect[0, 0], whiteR=rect[1, 0])
    pts_o, cnt_o = oracle.pointcloud_from_dense(xyz, valid, W, 400)
    assert n > 1000 and (cnt == cnt_o).all()
    assert (bits(sums[cnt > 0]) == bits(pts_o[cnt_o > 0])).all()


def test_reconstruct_ge_facade_with_auto_contrast(tmp_path, oracle):
    """Reconstruct::getParameters(autocontrast = true): loadCamImgs stretches every rectified image with
    Utilities::autoContrast before the decode (Duke/reconstruct.cpp:182-183).  Oracle-only (the reference's own
    autoContrast reads channels 1, 2 of a one-channel image: undefined behaviour)."""
    W, H = 320, 48
    stacks = synth.synth_gray(W, H, seed=64, noise_dn=2.0)
    stacks = (stacks.astype(np.float32) * 0.6 + 20).astype(np.uint8)      # a dim scan: the stretch matters
    nc = oracle.gray_num_bits(W)
    make_project(str(tmp_path), W, H, 0, stacks)
    sums, cnt, Q, m1, m2, _ = run_demo("ge", str(tmp_path), 0, W, 400, W, H, 40, 4, True, autocontrast=True)
    rect = np.stack([[oracle.auto_contrast(oracle.remap_linear(stacks[c, i], m1[c], m2[c])) for i in range(stacks.shape[1])]
                     for c in range(2)])
    dec = [oracle.gray_decode(rect[c], nc, 0, 40, 4, W, 400) for c in range(2)]
    xyz, valid, k, color, n = oracle.ge_triangulate(dec[0][0], dec[0][2], dec[1][0], dec[1][2], Q,
                                                    whiteL=rect[0, 0], whiteR=rect[1, 0])
    pts_o, cnt_o = oracle.pointcloud_from_dense(xyz, valid, W, 400)
    assert n > 300 and (cnt == cnt_o).all()
    assert (bits(sums[cnt > 0]) == bits(pts_o[cnt_o > 0])).all()
    # without the stretch the same dim scan decodes to a different cloud (the flag is not a no-op)
    sums0, cnt0, *_ = run_demo("ge", str(tmp_path), 0, W, 400, W, H, 40, 4, True, autocontrast=False)
    assert (cnt0 != cnt).any()


def test_auto_contrast_kernel_vs_oracle(oracle):
    import torch
    import slr_b200
    rng = np.random.default_rng(9)
    for W, H in ((320, 48), (37, 5)):                       # 128-bit path and the byte path (W*H % 16 != 0)
        eng = slr_b200.Engine(W, H, max_batch=1)
        imgs = np.stack([rng.integers(0, 256, (H, W)), rng.integers(90, 101, (H, W)), np.full((H, W), 77),
                         rng.integers(30, 200, (H, W)), np.clip(rng.normal(60, 10, (H, W)), 0, 255)]).astype(np.uint8)
        got = eng.auto_contrast(torch.from_numpy(imgs).cuda()).cpu().numpy()
        for q in range(len(imgs)):
            assert (got[q] == oracle.auto_contrast(imgs[q])).all(), (W, H, q)
        eng.close()


def test_reconstruct_gray_only_facade_end_to_end(tmp_path, oracle):
    W, H = 64, 48
    stacks = synth.synth_gray(W, H, seed=63, noise_dn=2.0, rows=True, integer_disparity=True)
    nc, nr = oracle.gray_num_bits(W), oracle.gray_num_bits(H)
    cams = make_project(str(tmp_path), W, H, 0, stacks)
    sums, cnt, _, _, _, _ = run_demo("gray", str(tmp_path), 0, W, H, W, H, 40, 3, False)
    dec = [oracle.gray_decode(stacks[c], nc, nr, 40, 3, W, H) for c in range(2)]
    s_o, c_o, n = oracle.gray_triangulate(dec[0][0], dec[0][1], dec[0][2], dec[1][0], dec[1][1], dec[1][2], W, H, cams)
    # facade PointCloudImage is h x w with element (j, i); the oracle reports cells in ac(i, j) = i*scan_h + j order
    c_img = c_o.reshape(W, H).T
    s_img = s_o.reshape(W, H, 3).transpose(1, 0, 2)
    assert (cnt == c_img).all()
    assert (bits(sums[cnt > 0]) == bits(s_img[c_img > 0])).all()


@pytest.mark.parametrize("color", [True, False])
@pytest.mark.parametrize("obj", [False, True])
def test_meshcreator_facade_writes_the_reference_file(tmp_path, oracle, color, obj):
    """MeshCreator (facade, GPU index passes + threaded text) writes byte for byte what the reference's
    exportPlyMesh / exportObjMesh write (Duke/meshcreator.cpp:16-166): golden fixture + the oracle on a larger cloud."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    duke = C.CDLL(os.path.join(ROOT, "structure-light-reconstructor_b200", "libduke_b200.so"))
    golden = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh.npz")))

    def export(pts, cnt, col, path):
        h, w = cnt.shape
        ci = None if col is None else np.ascontiguousarray(col, np.int32)
        nv, nf = C.c_ulonglong(0), C.c_ulonglong(0)
        rc = duke.duke_export_mesh(C.c_void_p(pts.ctypes.data), C.c_void_p(cnt.ctypes.data),
                                   C.c_void_p(ci.ctypes.data) if ci is not None else None, w, h, int(obj),
                                   str(path).encode(), C.byref(nv), C.byref(nf))
        assert rc == 0
        return nv.value, nf.value

    pts, cnt, col = cases.mesh_cloud(color=color)
    # the hook re-adds points one by one (count[q] addPoint calls): make the sums what that produces
    export(pts, cnt, col, tmp_path / "a.txt")
    want = golden[f"{'c' if color else 'n'}_{'obj' if obj else 'ply'}"].tobytes()
    assert open(tmp_path / "a.txt", "rb").read() == want

    rng = np.random.default_rng(77)
    h, w = 300, 417
    cnt = ((rng.random((h, w)) < 0.8) * 1).astype(np.uint8)
    pts = (rng.normal(0, 500, (h, w, 3)) * cnt[..., None]).astype(np.float32)
    col = rng.integers(0, 256, (h, w, 3)).astype(np.uint8) if color else None
    nv, nf = export(pts, cnt, col, tmp_path / "b.txt")
    oracle.export_mesh(pts, cnt, w, h, tmp_path / "o.txt", obj, None if col is None else col.astype(np.int32))
    assert open(tmp_path / "b.txt", "rb").read() == open(tmp_path / "o.txt", "rb").read()
    assert nv == int(cnt.sum()) and nf > 0
