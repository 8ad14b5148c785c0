"""CPU tests: the oracle against known answers derived by hand from the cited reference lines and
against a second, independent numpy restatement.  (The reference ships no fixtures: SURVEY.md §4.)"""
import math

import numpy as np
import pytest

import ref_np

F32 = np.float32


def test_gray_num_bits(oracle):
    # Duke/graycodes.cpp:24-25 (comments there: width 1280 -> 11, height 800 -> 10)
    for n, bits in [(640, 10), (1280, 11), (800, 10), (1024, 10), (2048, 11), (4096, 12), (480, 9), (3000, 12)]:
        assert oracle.gray_num_bits(n) == bits


def test_gray_num_imgs(oracle):
    # GRAY_ONLY 44, GRAY_EPI 24 at 1280x1024; config 1: 640 wide -> 22
    assert oracle.lib.orc_gray_num_imgs(1280, 1024, 1) == 24
    assert oracle.lib.orc_gray_num_imgs(1280, 1024, 0) == 44
    assert oracle.lib.orc_gray_num_imgs(640, 480, 1) == 22


def test_gray_to_dec_kat(oracle):
    # graycodes.cpp:116-128: prefix XOR from the MSB. 1011 -> prefix 1,1,0,1 -> 13
    assert oracle.gray_to_dec([1, 0, 1, 1]) == 13
    assert oracle.gray_to_dec([0, 0, 0, 0]) == 0
    assert oracle.gray_to_dec([1]) == 1
    assert oracle.gray_to_dec([1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]) == 0b10101010101
    for n in (4, 10, 11, 12):
        for j in range(0, 1 << n, 7):
            g = j ^ (j >> 1)
            bits = [(g >> (n - 1 - c)) & 1 for c in range(n)]
            assert oracle.gray_to_dec(bits) == j


@pytest.mark.parametrize("W,H,epi", [(640, 4, True), (1280, 3, True), (48, 40, False)])
def test_generate_gray_roundtrip(oracle, W, H, epi):
    g = oracle.generate_gray(W, H, epi)
    nc = oracle.gray_num_bits(W)
    assert (g[0] == 255).all() and (g[1] == 0).all()
    assert ((g[2::2].astype(int) + g[3::2].astype(int)) == 255).all()      # inverse images
    for j in range(W):
        bits = [1 if g[2 + 2 * c, 0, j] > g[3 + 2 * c, 0, j] else 0 for c in range(nc)]
        assert oracle.gray_to_dec(bits) == j
    if not epi:
        nr = oracle.gray_num_bits(H)
        for i in range(H):
            bits = [1 if g[2 + 2 * nc + 2 * c, i, 0] > g[3 + 2 * nc + 2 * c, i, 0] else 0 for c in range(nr)]
            assert oracle.gray_to_dec(bits) == i


def test_generate_mf_kat(oracle):
    m = oracle.generate_mf(1280, 2)
    assert (m[0] == 255).all() and (m[1] == 0).all()
    # multifrequency.cpp:27 at w = 0: 135+79*cos(PI*phi/2) -> 214, 134 (cos(1.5708)<0), 56, 135 (cos(4.7124)>0) (float -> uchar truncation)
    for f in range(3):
        assert [int(m[2 + 4 * f + s, 0, 0]) for s in range(4)] == [214, 134, 56, 135]
    assert m[2:].min() >= 56 and m[2:].max() <= 214
    # independent evaluation of the same expression
    for f, freq in enumerate((70, 64, 59)):
        for s in range(4):
            w = np.arange(1280, dtype=np.float64)
            arg = (3.1416 * 2 * w * freq / 1280 + 3.1416 * s / 2).astype(F32)
            exp = np.trunc(F32(135) + F32(79) * np.cos(arg)).astype(np.uint8)
            assert (m[2 + 4 * f + s, 1] == exp).mean() > 0.999   # numpy cos vs libm cosf may differ by an ulp


def test_shadow_mask_is_strict_greater(oracle):
    # mfreconstruct.cpp:201: whiteVal - blackVal > blackThreshold
    w = np.array([100, 100, 100, 20], np.uint8)
    b = np.array([60, 59, 61, 200], np.uint8)
    out = np.zeros(4, np.uint8)
    oracle.lib.orc_shadow_mask(w.ctypes.data, b.ctypes.data, 4, 40, out.ctypes.data)
    assert out.tolist() == [0, 1, 0, 0]


def test_wrapped_phase_branches_kat(oracle):
    PI = F32(3.1416)
    at = lambda q: F32(math.atan(q))
    cases = {
        (0, 5): F32(0.0),                         # :246 G4==G2 && G1>G3
        (0, -5): PI,                              # :248
        (7, 0): F32(3) * PI / F32(2),             # :250 (wrong quadrant, kept)
        (-7, 0): PI / F32(2),                     # :252
        (10, -3): F32(at(-3) + PI),               # :257 10/-3 = -3 (truncation toward zero)
        (-10, -3): F32(at(3) + PI),
        (10, 3): F32(at(3) + F32(2) * PI),        # :259
        (-10, 3): at(-3),                         # :261
        (2, 3): F32(F32(0.0) + F32(2) * PI),      # 2/3 = 0 -> atan(0) + 2PI
        (-2, 3): F32(0.0),                        # -2/3 = 0 (not -1)
        (255, 1): F32(at(255) + F32(2) * PI),
    }
    for (a, b), exp in cases.items():
        ok, P = oracle.wrapped_phase_strict(a, b)
        assert ok and P == exp, (a, b, P, exp)
    ok, _ = oracle.wrapped_phase_strict(0, 0)     # :254-255 degenerate -> pixel dropped
    assert not ok


def test_wrapped_phase_all_pairs_vs_numpy(oracle):
    a, b = np.meshgrid(np.arange(-255, 256), np.arange(-255, 256), indexing="ij")
    P_np, ok_np = ref_np.wrapped_phase(a, b)
    for ai in range(-255, 256, 1):
        for bi in (-255, -128, -3, -2, -1, 0, 1, 2, 3, 77, 255):
            ok, P = oracle.wrapped_phase_strict(ai, bi)
            assert ok == ok_np[ai + 255, bi + 255]
            if ok:
                assert P == P_np[ai + 255, bi + 255], (ai, bi)


def test_get_phase_vs_numpy_random(oracle):
    rng = np.random.default_rng(7)
    G = rng.integers(0, 256, size=(20000, 12))
    G[:200, 3] = G[:200, 1]           # exercise the a == 0 branches
    G[200:400, 0] = G[200:400, 2]     # and b == 0
    ph_np, ok_np = ref_np.get_phase(G)
    for n in range(G.shape[0]):
        ok, ph = oracle.get_phase_strict(G[n])
        assert ok == ok_np[n]
        if ok:
            assert ph == ph_np[n], (n, G[n], ph, ph_np[n])


def test_mf_decode_ideal_fringes_matches_survey_probe(oracle):
    # ideal 1280-px fringes -> 562 distinct strict phases in [-178, 475]
    m = oracle.generate_mf(1280, 1)
    stack = m.copy()
    stack[0] = 255
    ph, mk = oracle.mf_decode(stack, black_thr=40)
    assert mk.all()
    assert len(np.unique(ph)) == 562
    assert -179 < ph.min() < -178 and 474 < ph.max() < 475


def test_mf_decode_mask_and_degenerate(oracle):
    stack = np.zeros((14, 1, 4), np.uint8)
    stack[0] = 200
    stack[1] = [[20, 170, 20, 20]]            # pixel 1: 200-170 = 30 <= 40 -> shadow
    stack[2:] = 100                           # all four steps equal -> degenerate (G1==G3 && G4==G2)
    stack[2:, 0, 2] = np.arange(12) * 9 + 3   # pixel 2: a proper pixel
    ph, mk = oracle.mf_decode(stack, black_thr=40)
    assert mk.tolist() == [[0, 0, 1, 0]]
    assert np.isnan(ph[0, [0, 1, 3]]).all() and np.isfinite(ph[0, 2])


def test_undistort_vs_opencv(oracle):
    cv2 = pytest.importorskip("cv2")
    import slr_b200
    cams, _ = slr_b200.synthetic_rig(1280, 1024)
    cam = cams[0]
    K = np.array([[cam.fc[0], 0, cam.cc[0]], [0, cam.fc[1], cam.cc[1]], [0, 0, 1]], np.float64)
    D = np.array(list(cam.dist[:4]) + [0.0], np.float64)
    pts = np.array([[0, 0], [1279, 1023], [640, 512], [100, 900], [1200, 30]], np.float32)
    exp = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, P=K).reshape(-1, 2)
    for (x, y), e in zip(pts, exp):
        ox, oy = oracle.undistort_point(x, y, cam)
        assert abs(ox - e[0]) < 2e-2 and abs(oy - e[1]) < 2e-2      # same 5-iteration scheme (utilities.cpp:81-90)
    import dataclasses
    ox, oy = oracle.undistort_point(123.0, 456.0, dataclasses.replace(cam, dist=(0,) * 5))
    assert abs(ox - 123.0) < 1e-3 and abs(oy - 456.0) < 1e-3


def test_line_line_intersection_closed_form(oracle):
    # two rays through (1, 2, 10): utilities.cpp:399-425 returns the midpoint of the closest points
    p1, p2 = np.array([0, 0, 0], F32), np.array([5, 0, 0], F32)
    tgt = np.array([1, 2, 10], F32)
    v1 = (tgt - p1) / np.linalg.norm(tgt - p1)
    v2 = (tgt - p2) / np.linalg.norm(tgt - p2)
    ok, p = oracle.line_line_intersection(p1, v1, p2, v2)
    assert ok and np.allclose(p, tgt, atol=1e-4)
    ok, _ = oracle.line_line_intersection(p1, v1, p2, v1)            # parallel: |denom| < 0.1 -> reject
    assert not ok


def test_mf_triangulate_plane_closed_form(oracle):
    import slr_b200
    W, H, d = 64, 4, 5
    cams, Q = slr_b200.synthetic_rig(W, H, f=1000.0, tx=-100.0, distort=False)
    cams = [slr_b200.Camera(fc=(1000, 1000), cc=(32, 2)), slr_b200.Camera(fc=(1000, 1000), cc=(32, 2))]
    phR = np.tile(np.arange(W, dtype=F32) * F32(3.0), (H, 1))
    phL = np.full((H, W), np.nan, F32)
    phL[:, d:] = phR[:, :-d]                                         # left (j) sees right (j - d)
    mkL = np.isfinite(phL).astype(np.uint8)
    mkR = np.ones((H, W), np.uint8)
    xyz, valid, mk, n = oracle.mf_triangulate(np.nan_to_num(phL), mkL, phR, mkR, cams, Q)
    assert n == H * (W - d)
    assert (mk[:, d:] == np.arange(W - d)[None, :]).all() and (mk[:, :d] == -1).all()
    j, i = np.meshgrid(np.arange(W), np.arange(H))
    scale = -(-100.0) / d                                            # XYZ = (x-cx, y-cy, f) * (-Tx/d)
    exp = np.stack([(j - 32) * scale, (i - 2) * scale, np.full(j.shape, 1000.0 * scale)], -1)
    assert np.allclose(xyz[:, d:], exp[:, d:], rtol=1e-5, atol=1e-3)
    assert np.isnan(xyz[:, :d]).all() and not valid[:, :d].any()


def test_mf_match_is_first_k_within_tolerance(oracle):
    import slr_b200
    cams, Q = slr_b200.synthetic_rig(16, 1, distort=False)
    phR = np.array([[9.0, 5.05, 7.0, 5.0, 5.09, 5.2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], F32)
    mkR = np.ones((1, 16), np.uint8)
    mkR[0, 1] = 0                                                    # k=1 would match but carries no phase
    phL = np.full((1, 16), 5.0, F32)
    mkL = np.zeros((1, 16), np.uint8)
    mkL[0, 10] = 1
    _, valid, mk, n = oracle.mf_triangulate(phL, mkL, phR, mkR, cams, Q)
    assert n == 1 and mk[0, 10] == 3                                 # first valid k with |5-5| < 0.1
    phL[0, 10] = 5.15                                                # |5.15-5.09| < .1 at k=4 ; k=3: .15 no
    _, _, mk, _ = oracle.mf_triangulate(phL, mkL, phR, mkR, cams, Q)
    assert mk[0, 10] == 4
    phL[0, 10] = F32(5.1)                                            # fabs(5.1f - 5.0f) = 0.0999999 < 0.1 -> k=3
    _, _, mk, _ = oracle.mf_triangulate(phL, mkL, phR, mkR, cams, Q)
    assert mk[0, 10] == 3


def test_ge_kstart_chain(oracle):
    import slr_b200
    _, Q = slr_b200.synthetic_rig(8, 1, distort=False)
    colL = np.array([[5, 5, 7, 5, 9, 7, -1, 5]], np.int32)
    colR = np.array([[5, 7, 5, 7, 3, 3, 3, 3]], np.int32)
    mkL = (colL >= 0).astype(np.uint8)
    mkR = np.ones((1, 8), np.uint8)
    _, valid, mk, _, n = oracle.ge_triangulate(colL, mkL, colR, mkR, Q)
    # reconstruct.cpp:556-605: j0->k0 ; j1->k0 (kstart=k, not k+1) ; j2 (7)->k1 ; j3 (5): first k>=1 -> k2 ;
    # j4 (9): none, kstart stays 2 ; j5 (7): first k>=2 -> k3 ; j6 masked ; j7 (5): k>=3 -> none
    assert mk.tolist() == [[0, 0, 1, 2, -1, 3, -1, -1]]
    assert n == 5 and valid.sum() == 5


def test_ge_reprojection_uses_integer_pixels(oracle):
    import slr_b200
    _, Q = slr_b200.synthetic_rig(8, 2, f=500.0, tx=-50.0, distort=False)
    colL = np.full((2, 8), -1, np.int32)
    colR = np.full((2, 8), -1, np.int32)
    colL[1, 6] = 3
    colR[1, 2] = 3
    xyz, valid, mk, _, n = oracle.ge_triangulate(colL, (colL >= 0).astype(np.uint8), colR,
                                                 (colR >= 0).astype(np.uint8), Q)
    assert n == 1 and mk[1, 6] == 2
    s = 50.0 / 4                                                      # -Tx / (j - k)
    assert np.allclose(xyz[1, 6], [(6 - 4) * s, (1 - 1) * s, 500.0 * s], rtol=1e-6)


def test_pointcloud_drop_rule_F7(oracle):
    # addPoint(row, col): dropped when row >= scan_w or col >= scan_h (pointcloudimage.cpp:88)
    W, H, scan_w, scan_h = 8, 4, 3, 6
    xyz = np.arange(H * W * 3, dtype=F32).reshape(H, W, 3)
    valid = np.ones((H, W), np.uint8)
    pts, cnt = oracle.pointcloud_from_dense(xyz, valid, scan_w, scan_h)
    assert cnt.shape == (scan_h, scan_w)
    for i in range(H):
        for j in range(W):
            if i < scan_w and j < scan_h:
                assert cnt[j, i] == 1 and (pts[j, i] == xyz[i, j]).all()   # stored transposed
    assert cnt.sum() == min(H, scan_w) * min(W, scan_h)


def test_gray_decode_kat(oracle):
    import slr_b200
    W, H = 32, 2
    g = slr_b200.generate_gray_patterns(W, H, True).astype(np.int32)
    nb = oracle.gray_num_bits(W)
    stack = np.clip(g * 200 // 255 + 20, 0, 255).astype(np.uint8)     # pattern 220 / 20
    stack[0], stack[1] = 220, 20
    col, _, mk = oracle.gray_decode(stack, nb, black_thr=40, white_thr=5, scan_w=W)
    assert mk.all() and (col == np.arange(W)[None, :]).all()
    stack2 = stack.copy()
    stack2[4, 0, 7] = stack2[5, 0, 7]                                 # |v1-v2| = 0 < whiteThreshold -> error
    stack2[1, 1, 3] = 200                                             # shadow
    col2, _, mk2 = oracle.gray_decode(stack2, nb, black_thr=40, white_thr=5, scan_w=W)
    assert mk2[0, 7] == 0 and col2[0, 7] == -1 and mk2[1, 3] == 0
    col3, _, mk3 = oracle.gray_decode(stack, nb, black_thr=40, white_thr=0, scan_w=20)
    assert mk3[0, :21].all() and not mk3[0, 21:].any()                # xDec > scan_w is strict (:403)


def test_auto_contrast_known_answers(oracle):
    """Utilities::autoContrast (Duke/utilities.cpp:340-355), channel 0, by hand: min = 20 + 12.75 = 32.75,
    a = 255 / (220 - 32.75) = 1.36181...; v = 20 -> 0 (saturated), v = 33 -> round(0.25) = 0 -> 0,
    v = 34 -> round(1.25) = 1 -> round(1.3618) = 1, v = 120 -> round(87.25) = 87 -> round(118.48) = 118,
    v = 220 -> round(187.25) = 187 -> round(254.66) = 255."""
    img = np.array([[20, 33, 34, 120, 220]], np.uint8)
    assert oracle.auto_contrast(img).tolist() == [[0, 0, 1, 118, 255]]
    # a flat image: max - (min + 12.75) < 0, every pixel saturates to 0 in the subtraction
    assert (oracle.auto_contrast(np.full((3, 4), 99, np.uint8)) == 0).all()
