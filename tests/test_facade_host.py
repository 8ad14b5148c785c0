"""CPU tests of the facade's host logic (no GPU): PNG/PGM IO, the cv::stereoRectify / initUndistortRectifyMap
restatement against the cv2 of this image (in its MODERN variant; the product follows OpenCV 2.4.9, which differs in
two documented details), stereoRect's host remap against cv2.remap, VirtualCamera's loaders."""
import ctypes as C
import os
import struct
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "structure-light-reconstructor_b200", "libduke_b200.so")


@pytest.fixture(scope="module")
def duke():
    if not os.path.exists(LIB):
        pytest.fail(f"{LIB} is not built (python -c 'import __graft_entry__ as g; g.build()')")
    return C.CDLL(LIB)


def f32(a):
    return np.asarray(a, np.float64).astype(np.float32).astype(np.float64)


def ptr(a):
    return C.c_void_p(a.ctypes.data)


def rig(seed):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    M1 = f32([[2400 + rng.normal(0, 9), 0, 640 + rng.normal(0, 5)], [0, 2402 + rng.normal(0, 9), 512 + rng.normal(0, 5)], [0, 0, 1]])
    M2 = f32([[2395, 0, 645 + rng.normal(0, 5)], [0, 2397, 508 + rng.normal(0, 5)], [0, 0, 1]])
    D1 = f32([-0.12, -0.08, 0.0007, -0.0004, 0.01])
    D2 = f32([0.10, 0.05, -0.0005, 0.0006, -0.02])
    R, _ = cv2.Rodrigues(np.array([0.01 * rng.normal(), 0.15, -0.01]))
    return M1, D1, M2, D2, f32(R), f32([-200, 2 * rng.normal(), 5])


def run_rectify(duke, rigv, W, H, variant):
    inp = np.concatenate([np.asarray(a, np.float64).ravel() for a in rigv])
    out = np.zeros(9 + 9 + 12 + 12 + 16)
    duke.duke_stereo_rectify(ptr(inp), W, H, variant, ptr(out))
    return out[:9].reshape(3, 3), out[9:18].reshape(3, 3), out[18:30].reshape(3, 4), out[30:42].reshape(3, 4), out[42:].reshape(4, 4)


def build_map(duke, M, D, R, P, W, H):
    a = np.zeros((H, W, 2), np.int16)
    b = np.zeros((H, W), np.uint16)
    args = [np.ascontiguousarray(x, np.float64) for x in (M, D, R, P)]
    duke.duke_init_undistort_rectify_map(*[ptr(x) for x in args], W, H, ptr(a), ptr(b))
    return a, b


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_stereo_rectify_matches_opencv_modern_variant(duke, seed):
    cv2 = pytest.importorskip("cv2")
    W, H = 1280, 1024
    rv = rig(seed)
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(rv[0], rv[1], rv[2], rv[3], (W, H), rv[4], rv[5], flags=0, alpha=-1)
    mine = run_rectify(duke, rv, W, H, 1)
    for a, b in zip((R1, R2, P1, P2, Q), mine):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-9)
    m1, m2 = cv2.initUndistortRectifyMap(rv[0], rv[1], R1, P1, (W, H), cv2.CV_16SC2)
    a, b = build_map(duke, rv[0], rv[1], mine[0], mine[2], W, H)
    assert (a == m1).all() and (b == m2).all()


def test_stereo_rectify_cv249_variant_differs_only_as_documented(duke):
    """2.4.9: fc_new = min over cameras of fy, shrunk by 1 + k1 (nx^2+ny^2)/(4 fy^2) when k1 < 0; principal point from
    (nx-1)/2 in integer division.  R1/R2 are unaffected; Q keeps the stereoRectify layout."""
    W, H = 1280, 1024
    rv = rig(4)
    old = run_rectify(duke, rv, W, H, 0)
    new = run_rectify(duke, rv, W, H, 1)
    assert np.allclose(old[0], new[0]) and np.allclose(old[1], new[1])
    fy1, fy2, k1 = rv[0][1, 1], rv[2][1, 1], rv[1][0]
    fc = min(fy1 * (1 + k1 * (W * W + H * H) / (4 * fy1 * fy1)), fy2)
    assert abs(old[2][0, 0] - fc) < 1e-9 and abs(new[2][0, 0] - (fy1 + fy2) / 2) < 1e-9
    Q = old[4]
    assert Q[0, 0] == 1 and Q[1, 1] == 1 and Q[2, 3] == old[2][0, 0] and Q[3, 2] > 0 and Q[0, 3] == -old[2][0, 2]


def _png_with_all_filters(rgb):
    def filt(y, ft):
        cur = rgb[y].astype(int).ravel()
        up = rgb[y - 1].astype(int).ravel() if y else np.zeros_like(cur)
        out = []
        for x in range(cur.size):
            a = cur[x - 3] if x >= 3 else 0
            b = up[x]
            c = up[x - 3] if x >= 3 else 0
            pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
            pred = [0, a, b, (a + b) // 2, a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)][ft]
            out.append((cur[x] - pred) & 255)
        return bytes([ft]) + bytes(out)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    h, w = rgb.shape[:2]
    raw = b"".join(filt(y, y % 5) for y in range(h))
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def test_png_pgm_roundtrip_and_filters(duke, tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    for name, writer in (("a.png", duke.duke_write_png_gray), ("a.pgm", duke.duke_write_pgm)):
        p = str(tmp_path / name).encode()
        assert writer(p, ptr(img), 53, 37) == 0
        w, h = C.c_int(0), C.c_int(0)
        out = np.zeros(37 * 53, np.uint8)
        assert duke.duke_read_gray_image(p, C.byref(w), C.byref(h), ptr(out), out.size) == 0
        assert (w.value, h.value) == (53, 37) and (out.reshape(37, 53) == img).all()
    # an RGB PNG using every filter type: decoded and reduced to gray like cv::imread(path, 0)
    rgb = rng.integers(0, 256, (9, 11, 3)).astype(np.uint8)
    p = tmp_path / "rgb.png"
    p.write_bytes(_png_with_all_filters(rgb))
    w, h = C.c_int(0), C.c_int(0)
    out = np.zeros(99, np.uint8)
    assert duke.duke_read_gray_image(str(p).encode(), C.byref(w), C.byref(h), ptr(out), 99) == 0
    exp = (rgb[..., 0].astype(int) * 4899 + rgb[..., 1].astype(int) * 9617 + rgb[..., 2].astype(int) * 1868 + 8192) >> 14
    assert (out.reshape(9, 11) == exp).all()
    cv2 = pytest.importorskip("cv2")
    assert np.abs(cv2.imread(str(p), 0).astype(int) - out.reshape(9, 11)).max() <= 1
    assert duke.duke_read_gray_image(str(tmp_path / "missing.png").encode(), C.byref(w), C.byref(h), None, 0) == -1


def _inflate(duke, comp, n):
    out = np.full(n + 8, 0xAB, np.uint8)
    src = np.frombuffer(comp, np.uint8)
    rc = duke.duke_zlib_inflate(ptr(src), C.c_size_t(len(comp)), ptr(out), C.c_size_t(n))
    assert (out[n:] == 0xAB).all(), "wrote past the end of the output"
    return rc, out[:n]


def test_own_inflate_equals_zlib_on_every_level_and_strategy(duke):
    """facade/inflate.cpp (the PNG ingest path's entropy decoder) against zlib streams of every compression level,
    strategy and memory level — stored, fixed and dynamic blocks, long codes (second-level tables), long distances —
    and against damaged input, which must be rejected, not trusted."""
    rng = np.random.default_rng(3)
    n_cases = 0
    for kind in range(7):
        for sz in (0, 1, 7, 300, 5000, 70000, 300000):
            if kind == 0:
                raw = rng.integers(0, 256, sz, dtype=np.uint8)                                  # incompressible
            elif kind == 1:
                raw = np.zeros(sz, np.uint8)
            elif kind == 2:                                                                     # noisy fringe, Sub-filtered
                x = np.arange(sz)
                v = (128 + 60 * np.sin(x * 0.05) + rng.integers(-3, 4, sz)).astype(np.int16)
                raw = (np.diff(v, prepend=0) & 255).astype(np.uint8)
            elif kind == 3:
                raw = np.frombuffer((b"abcabcabdabc" * (sz // 12 + 1))[:sz], np.uint8)          # short distances
            elif kind == 4:
                raw = (np.arange(sz) >> 8).astype(np.uint8)                                     # long runs
            elif kind == 5:
                raw = np.where(rng.random(sz) < 0.03, rng.integers(0, 256, sz), 7).astype(np.uint8)
            else:                                                                               # alphabet grows: long codes
                raw = (rng.integers(0, 1 << 30, sz) % (1 + (np.arange(sz) >> 10))).astype(np.uint8)
            for level in (0, 1, 6, 9):
                for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                    for mem in (1, 9):
                        c = zlib.compressobj(level, zlib.DEFLATED, 15, mem, strategy)
                        comp = c.compress(raw.tobytes()) + c.flush()
                        rc, out = _inflate(duke, comp, sz)
                        assert rc == 0 and (out == raw).all(), (kind, sz, level, strategy, mem)
                        n_cases += 1
            if sz >= 300:
                comp = zlib.compress(raw.tobytes(), 6)
                assert _inflate(duke, comp[:len(comp) // 2], sz)[0] != 0          # truncated
                assert _inflate(duke, comp, sz - 1)[0] != 0                       # stream longer than the output
                assert _inflate(duke, comp, sz + 1)[0] != 0                       # stream shorter than the output
                bad = bytearray(comp)
                bad[-1] ^= 1                                                      # Adler-32
                assert _inflate(duke, bytes(bad), sz)[0] != 0
                for pos in rng.integers(2, len(comp) - 4, 40):                    # bit rot: rejected or (rarely) same length,
                    bad = bytearray(comp)                                         # never a crash or an overrun
                    bad[pos] ^= 1 << int(rng.integers(0, 8))
                    _inflate(duke, bytes(bad), sz)
    assert n_cases == 7 * 7 * 4 * 5 * 2
    assert _inflate(duke, b"\x78\x9c", 0)[0] != 0 and _inflate(duke, b"", 0)[0] != 0


def write_mat(path, m):
    with open(path, "w") as f:
        for r in np.atleast_2d(m):
            f.write("\t".join(f"{v:.9g}" for v in r) + "\t\n")


def test_stereorect_loads_project_and_remaps_like_opencv(duke, tmp_path):
    cv2 = pytest.importorskip("cv2")
    W, H = 320, 200
    M1 = [[600, 0, 162.25], [0, 601, 98.5], [0, 0, 1]]
    M2 = [[598, 0, 158.25], [0, 600.5, 101.25], [0, 0, 1]]
    D1, D2 = [-0.11, 0.07, 0.0006, -0.0003, 0.0], [-0.09, 0.04, -0.0004, 0.0005, 0.0]
    R = [[0.99985, 0.002, 0.0172], [-0.0021, 0.999995, 0.004], [-0.0172, -0.004, 0.99984]]
    T = [-80.0, 0.4, 1.2]
    for side, M, D in (("left", M1, D1), ("right", M2, D2)):
        os.makedirs(tmp_path / "calib" / side)
        write_mat(tmp_path / "calib" / side / "cam_stereo.txt", M)
        write_mat(tmp_path / "calib" / side / "distortion_stereo.txt", np.array(D)[:, None])
    write_mat(tmp_path / "calib" / "R_stereo.txt", R)
    write_mat(tmp_path / "calib" / "T_stereo.txt", np.array(T)[:, None])
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (H, W)).astype(np.uint8)
    Q = np.zeros(16)
    out = img.copy()
    assert duke.duke_stereorect_probe(str(tmp_path).encode(), W, H, ptr(Q), ptr(out), 1) == 0
    rv = [f32(x) for x in (M1, D1, M2, D2, R, T)]
    R1, R2, P1, P2, Q2 = run_rectify(duke, rv, W, H, 0)
    assert np.array_equal(Q.reshape(4, 4), Q2)
    a, b = build_map(duke, rv[0], rv[1], R1, P1, W, H)
    assert (cv2.remap(img, a, b, cv2.INTER_LINEAR) == out).all()
    assert duke.duke_stereorect_probe(str(tmp_path / "nope").encode(), W, H, ptr(Q), ptr(out), 1) == -1


def test_virtualcamera_loader(duke, tmp_path):
    write_mat(tmp_path / "cam_matrix.txt", [[2400.123456789, 0, 640.5], [0, 2401.25, 511.75], [0, 0, 1]])
    v = (C.c_float * 4)()
    assert duke.duke_load_camera_matrix(str(tmp_path / "cam_matrix.txt").encode(), v) == 1
    assert list(v) == [np.float32(2400.123456789), np.float32(2401.25), np.float32(640.5), np.float32(511.75)]
    assert duke.duke_load_camera_matrix(str(tmp_path / "missing.txt").encode(), v) == 0


@pytest.mark.parametrize("color", [True, False])
@pytest.mark.parametrize("obj", [False, True])
def test_mesh_text_stage_writes_the_reference_bytes(duke, tmp_path, color, obj):
    """MeshCreator's text stage (threaded std::to_chars formatting) on the oracle's index arrays == the reference's
    exportPlyMesh / exportObjMesh bytes (golden fixture made by the reference's own MeshCreator).  The index passes
    themselves run on the GPU and are checked in the -m gpu suite."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import oracle_lib
    orc = oracle_lib.load()
    golden = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh.npz")))
    pts, cnt, col = cases.mesh_cloud(color=color)
    h, w = cnt.shape
    vert, src, faces = orc.mesh_index(pts, cnt, w, h, 1 if obj else 0)
    vert, src, faces = np.ascontiguousarray(vert), np.ascontiguousarray(src), np.ascontiguousarray(faces)
    ci = None if col is None else np.ascontiguousarray(col, np.int32)
    path = tmp_path / "m.txt"
    rc = duke.duke_write_mesh_text(ptr(pts), ptr(cnt), ptr(ci) if ci is not None else None, w, h, int(obj), ptr(vert), ptr(src),
                                   ptr(faces), C.c_ulonglong(len(vert)), C.c_ulonglong(len(faces)), str(path).encode())
    assert rc == 0
    assert open(path, "rb").read() == golden[f"{'c' if color else 'n'}_{'obj' if obj else 'ply'}"].tobytes()
    # a larger cloud: enough items for the multi-threaded formatter (>= 4096), against the oracle's writer
    rng = np.random.default_rng(5)
    h, w = 120, 97
    cnt = (rng.random((h, w)) < 0.8).astype(np.uint8)
    pts = (rng.normal(0, 400, (h, w, 3)) * cnt[..., None]).astype(np.float32)
    col = rng.integers(0, 256, (h, w, 3)).astype(np.uint8) if color else None
    vert, src, faces = orc.mesh_index(pts, cnt, w, h, 1 if obj else 0)
    vert, src, faces = np.ascontiguousarray(vert), np.ascontiguousarray(src), np.ascontiguousarray(faces)
    ci = None if col is None else np.ascontiguousarray(col, np.int32)
    rc = duke.duke_write_mesh_text(ptr(pts), ptr(cnt), ptr(ci) if ci is not None else None, w, h, int(obj), ptr(vert), ptr(src),
                                   ptr(faces), C.c_ulonglong(len(vert)), C.c_ulonglong(len(faces)), str(tmp_path / "b.txt").encode())
    assert rc == 0 and len(vert) >= 4096
    orc.export_mesh(pts, cnt, w, h, tmp_path / "o.txt", obj, ci)
    assert open(tmp_path / "b.txt", "rb").read() == open(tmp_path / "o.txt", "rb").read()


@pytest.mark.parametrize("color", [True, False])
@pytest.mark.parametrize("export_off", [True, False])
def test_pointcloudimage_export_xyz_writes_the_reference_bytes(duke, tmp_path, color, export_off):
    """PointCloudImage::exportXYZ (Duke/pointcloudimage.cpp:99-122): the facade's accumulator writes the bytes the
    reference's own class writes (fixture made by oracle/_ref; re-derived live where /root/reference is mounted)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import ref_lib
    golden = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh.npz")))
    pts, cnt, col = cases.mesh_cloud(color=color)
    h, w = cnt.shape
    ci = None if col is None else np.ascontiguousarray(col, np.int32)
    path = tmp_path / "c.xyz"
    rc = duke.duke_export_xyz(ptr(pts), ptr(cnt), ptr(ci) if ci is not None else None, w, h, int(export_off), 1, str(path).encode())
    assert rc == 0
    want = golden[f"{'c' if color else 'n'}_xyz_{'off' if export_off else 'on'}"].tobytes()
    assert open(path, "rb").read() == want
    if ref_lib.available():
        ref_lib.load().export_xyz(pts, cnt, w, h, tmp_path / "r.xyz", export_off, True, col)
        assert open(tmp_path / "r.xyz", "rb").read() == want


def test_ingest_host_half_routes_every_encoding(duke, tmp_path):
    """facade/ingest.cpp's per-image work (no GPU): OpenCV-style Sub-filtered PNGs stay scanlines for the GPU unfilter,
    files with Up rows are flagged, Average / Paeth rows are unfiltered on the host, PGM / colour PNG arrive as pixels,
    wrong sizes / damaged / missing files are refused."""
    from slr_b200 import synth
    rng = np.random.default_rng(4)
    W, H = 64, 21
    img = rng.integers(0, 256, (H, W)).astype(np.uint8)
    out = np.zeros(H * (W + 1), np.uint8)

    def decode(base, w=W, h=H):
        out[:] = 0xEE
        return duke.duke_decode_scan_image(str(tmp_path / base).encode(), b".png", w, h, ptr(out))

    synth.write_png_opencv_style(str(tmp_path / "sub.png"), img)
    assert decode("sub") == 1
    rows = out.reshape(H, W + 1)
    assert (rows[:, 0] == 1).all() and (np.cumsum(rows[:, 1:], axis=1, dtype=np.uint8) == img).all()

    def write_filtered(name, types):
        a = img.astype(np.int32)
        raw = b""
        for y in range(H):
            t = types[y]
            left = np.concatenate([[0], a[y, :-1]])
            up = a[y - 1] if y else np.zeros(W, np.int32)
            ul = np.concatenate([[0], up[:-1]])
            pa, pb, pc = np.abs(up - ul), np.abs(left - ul), np.abs(left + up - 2 * ul)
            pred = [0 * left, left, up, (left + up) >> 1, np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))][t]
            raw += bytes([t]) + ((a[y] - pred) & 255).astype(np.uint8).tobytes()
        z = zlib.compress(raw, 6)

        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
        (tmp_path / name).write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 0, 0, 0, 0)) +
                                      chunk(b"IDAT", z[:7]) + chunk(b"IDAT", z[7:]) + chunk(b"IEND", b""))

    write_filtered("up.png", [y % 3 for y in range(H)])
    assert decode("up") == 2 and (out.reshape(H, W + 1)[:, 0] == [y % 3 for y in range(H)]).all()
    write_filtered("paeth.png", [y % 5 for y in range(H)])
    assert decode("paeth") == 3 and (out[:H * W].reshape(H, W) == img).all()
    assert duke.duke_write_pgm(str(tmp_path / "only_pgm.pgm").encode(), ptr(img), W, H) == 0
    assert decode("only_pgm") == 3 and (out[:H * W].reshape(H, W) == img).all()
    rgb = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    (tmp_path / "rgb.png").write_bytes(_png_with_all_filters(rgb))
    assert decode("rgb") == 3
    exp = (rgb[..., 0].astype(int) * 4899 + rgb[..., 1].astype(int) * 9617 + rgb[..., 2].astype(int) * 1868 + 8192) >> 14
    assert (out[:H * W].reshape(H, W) == exp).all()
    assert decode("sub", W + 4, H) == 0 and decode("missing") == 0
    blob = bytearray((tmp_path / "sub.png").read_bytes())
    blob[60] ^= 0x10
    (tmp_path / "bad.png").write_bytes(bytes(blob))
    assert decode("bad") == 0
    (tmp_path / "cut.png").write_bytes(bytes(blob[:len(blob) // 2]))
    assert decode("cut") == 0
