"""CPU tests that PIN the oracle against the reference's own code.

tests/golden/ref_*.npz hold the outputs of the reference's sources (Duke/*.cpp compiled unmodified against
oracle/ref_shim/, see oracle/Makefile `ref`) on the seeded cases of tests/golden/cases.py.  The oracle must
reproduce them bit for bit.  Where oracle/_ref/libref.so exists (the build container) the fixtures are also
re-derived live, so a stale fixture cannot hide a regression."""
import os

import numpy as np
import pytest

import ref_lib

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases  # noqa: E402


def golden(name):
    return dict(np.load(os.path.join(HERE, "golden", f"ref_{name}.npz")))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_patterns_match_reference(oracle):
    g = golden("patterns")
    assert (oracle.generate_gray(80, 4, True) == g["gray_epi_80x4"]).all()
    assert (oracle.generate_gray(48, 40, False) == g["gray_full_48x40"]).all()
    assert (oracle.generate_mf(1280, 2) == g["mf_1280x2"]).all()
    for (w, h, e), (n, nc, nr) in zip([(1280, 1024, 1), (1280, 1024, 0), (640, 480, 1), (2048, 1536, 1), (800, 600, 0)],
                                      g["layouts"]):
        assert oracle.lib.orc_gray_num_imgs(w, h, e) == n
        assert oracle.gray_num_bits(w) == nc and oracle.gray_num_bits(h) == nr
    got = np.array([oracle.gray_to_dec([(v >> (10 - c)) & 1 for c in range(11)]) for v in range(2048)])
    assert (got == g["g2d"]).all()


@pytest.mark.parametrize("case", ["mf_pairs", "mf_scene"])
def test_mf_decode_matches_reference(oracle, case):
    g = golden(case)
    stacks = g["stack"] if g["stack"].ndim == 4 else g["stack"][None]
    for cam in range(stacks.shape[0]):
        ph_r = g["phase"][cam] if g["phase"].ndim == 3 else g["phase"]
        has_r = g["has"][cam] if g["has"].ndim == 3 else g["has"]
        mk_r = g["mask"][cam] if g["mask"].ndim == 3 else g["mask"]
        ph_o, mk_o = oracle.mf_decode(stacks[cam], black_thr=40)
        # final mask identical (shadow test + degenerate branch)
        assert (mk_o == mk_r).all()
        # the reference also pushes a phase for degenerate pixels (has=1, mask=0): undefined value (F4), excluded
        deg = (has_r == 1) & (mk_r == 0)
        assert ((has_r == 1) == ((mk_o == 1) | deg)).all()
        ok = mk_r == 1
        assert (bits(ph_o[ok]) == bits(ph_r[ok])).all()
    if case == "mf_pairs":
        assert ((g["has"] == 1) & (g["mask"] == 0)).sum() > 0     # the sweep does contain degenerate pixels


def _dense_from_cloud(points, count, W, H):
    """PointCloudImage stores addPoint(row, col) at Mat(col, row): back to [H][W]"""
    return points.transpose(1, 0, 2)[:H, :W], count.T[:H, :W]


@pytest.mark.parametrize("case,rigid", [("mf_tri", None), ("mf_tri_rigid", cases.RIGID)])
def test_mf_triangulation_matches_reference(oracle, case, rigid):
    sc = golden("mf_scene")
    g = golden(case)
    H, W = sc["stack"].shape[2:]
    cams, Q = cases.rig(W, H)
    xyz, valid, k, n = oracle.mf_triangulate(sc["phase"][0], sc["mask"][0], sc["phase"][1], sc["mask"][1], cams, Q, rigid)
    pts_r, cnt_r = _dense_from_cloud(g["points"], g["count"], W, H)
    assert (cnt_r == valid).all() and n == int(cnt_r.sum()) and n > 100
    assert (bits(xyz[valid == 1]) == bits(pts_r[valid == 1])).all()
    # and the oracle's PointCloudImage adapter reproduces the container itself
    pts_o, cnt_o = oracle.pointcloud_from_dense(xyz, valid, g["count"].shape[1], g["count"].shape[0])
    assert (cnt_o == g["count"]).all() and (bits(pts_o[cnt_o > 0]) == bits(g["points"][cnt_o > 0])).all()


def test_pointcloud_drop_rule_matches_reference(oracle):
    sc = golden("mf_scene")
    g = golden("mf_tri_drop")                       # PointCloudImage(scan_w=6, scan_h=100): rows >= 6, cols >= 100 dropped
    H, W = sc["stack"].shape[2:]
    cams, Q = cases.rig(W, H)
    xyz, valid, _, _ = oracle.mf_triangulate(sc["phase"][0], sc["mask"][0], sc["phase"][1], sc["mask"][1], cams, Q)
    pts_o, cnt_o = oracle.pointcloud_from_dense(xyz, valid, 6, 100)
    assert (cnt_o == g["count"]).all() and cnt_o.sum() > 0
    assert (bits(pts_o[cnt_o > 0]) == bits(g["points"][cnt_o > 0])).all()


def test_gray_epi_decode_and_triangulation_match_reference(oracle):
    g = golden("ge_decode")
    H, W = g["stack"].shape[2:]
    nc = oracle.gray_num_bits(W)
    cols, mks = [], []
    for cam in range(2):
        c, _, m = oracle.gray_decode(g["stack"][cam], nc, 0, 40, 5, W, H)
        assert (m == g["mask"][cam]).all() and (c == g["col"][cam]).all()
        cols.append(c)
        mks.append(m)
    t = golden("ge_tri")
    _, Q = cases.rig(W, H)
    xyz, valid, k, color, n = oracle.ge_triangulate(cols[0], mks[0], cols[1], mks[1], Q, cases.RIGID,
                                                    g["stack"][0, 0], g["stack"][1, 0])
    pts_r, cnt_r = _dense_from_cloud(t["points"], t["count"], W, H)
    assert (cnt_r == valid).all() and n > 100
    assert (bits(xyz[valid == 1]) == bits(pts_r[valid == 1])).all()
    assert (t["color"].T[:H, :W][valid == 1] == color[valid == 1]).all()


def test_gray_only_decode_and_bucket_triangulation_match_reference(oracle):
    g = golden("go_decode")
    H, W = g["stack"].shape[2:]
    nc, nr = oracle.gray_num_bits(W), oracle.gray_num_bits(H)
    dec = []
    for cam in range(2):
        c, r, m = oracle.gray_decode(g["stack"][cam], nc, nr, 40, 3, W, H)
        assert (m == g["mask"][cam]).all()
        # the reference files a pixel only when its cell index is inside the table; compare on those
        inside = g["col"][cam] >= 0
        assert (c[inside] == g["col"][cam][inside]).all() and (r[inside] == g["row"][cam][inside]).all()
        assert (m[inside] == 1).all()
        dec.append((c, r, m))
    t = golden("go_tri")
    ssum, cnt, n = oracle.gray_triangulate(dec[0][0], dec[0][1], dec[0][2], dec[1][0], dec[1][1], dec[1][2], W, H,
                                           cases.gray_only_rig(W, H))
    assert (cnt == t["cnt"]).all() and n > 20
    assert (bits(ssum[cnt > 0]) == bits(t["sum"][cnt > 0])).all()


def test_geometry_helpers_match_reference(oracle):
    g = golden("helpers")
    cams, _ = cases.rig(1280, 1024)
    for c in range(2):
        got = np.array([oracle.undistort_point(x, y, cams[c]) for (x, y) in g["px"]], np.float32)
        assert (bits(got) == bits(g["undistort"][c])).all()
    camsg = cases.gray_only_rig(48, 40)
    import ctypes as C
    import oracle_lib
    for i, v in enumerate(g["vecs"]):
        ok, p = oracle.line_line_intersection(v[0] * 50, v[1] / np.linalg.norm(v[1]), v[2] * 50, v[3] / np.linalg.norm(v[3]))
        assert ok == bool(g["ll_ok"][i])
        if ok:
            assert (bits(p) == bits(g["ll_p"][i])).all()
        q = np.ascontiguousarray(v[0] * 10, np.float32)
        cam = oracle_lib._cam(camsg[i % 2])
        oracle.lib.orc_cam2world(C.byref(cam), C.c_void_p(q.ctypes.data))
        assert (bits(q) == bits(g["cam2world"][i])).all()
        w = np.ascontiguousarray(v[1] * (1e-7 if i % 9 == 0 else 3.0), np.float32)
        oracle.lib.orc_normalize(C.c_void_p(w.ctypes.data))
        assert (bits(w) == bits(g["normalize"][i])).all()
    assert g["ll_ok"].sum() > 5 and (~g["ll_ok"].astype(bool)).sum() >= 2      # both outcomes of |denom| < 0.1 occur
    # PointCloudImage::addPoint sequence with a cell that receives 300 points (u8 count wraps, sum resets)
    pts = np.empty((4, 6, 3), np.float32)
    cnt = np.empty((4, 6), np.uint8)
    iw, jh, pp = (np.ascontiguousarray(g[k]) for k in ("pc_iw", "pc_jh", "pc_pts"))
    oracle.lib.orc_pointcloud_add(6, 4, C.c_void_p(iw.ctypes.data), C.c_void_p(jh.ctypes.data), C.c_void_p(pp.ctypes.data),
                                  len(iw), C.c_void_p(pts.ctypes.data), C.c_void_p(cnt.ctypes.data))
    assert (cnt == g["pc_cnt"]).all() and (bits(pts[cnt > 0]) == bits(g["pc_sum"][cnt > 0])).all()
    assert g["pc_cnt"][3, 2] == (300 + int(((g["pc_iw"][300:] == 2) & (g["pc_jh"][300:] == 3)).sum())) % 256


@pytest.mark.parametrize("tag,color", [("c", True), ("n", False)])
@pytest.mark.parametrize("obj", [False, True])
def test_mesh_export_matches_reference(oracle, tmp_path, tag, color, obj):
    """MeshCreator::exportPlyMesh / exportObjMesh (Duke/meshcreator.cpp:16-166): the oracle writes the same bytes."""
    g = golden("mesh")
    pts, cnt, col = cases.mesh_cloud(color=color)
    h, w = cnt.shape
    path = tmp_path / "m.txt"
    oracle.export_mesh(pts, cnt, w, h, path, obj, None if col is None else col.astype(np.int32))
    want = g[f"{tag}_{'obj' if obj else 'ply'}"].tobytes()
    assert open(path, "rb").read() == want
    if ref_lib.available():                       # live: the reference's own MeshCreator, now
        rp = tmp_path / "r.txt"
        ref_lib.load().export_mesh(pts, cnt, w, h, rp, obj, col)
        assert open(rp, "rb").read() == want
    # the index arrays agree with the text (vertex count in the header / number of "v" lines)
    vert, src, faces = oracle.mesh_index(pts, cnt, w, h, 1 if obj else 0)
    text = want.decode()
    if obj:
        assert text.count("\nv ") + text.startswith("v ") == len(vert) and text.count("\nf ") == len(faces)
    else:
        assert f"element vertex {len(vert)}\n" in text and f"element face {len(faces)}\n" in text
    assert len(vert) == int((cnt > 0).sum())


@pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libref.so only exists where /root/reference is mounted")
def test_fixtures_are_what_the_reference_computes_now():
    """Re-run the reference's code and require the committed fixtures to be current."""
    r = ref_lib.load()
    g = golden("mf_pairs")
    ph, has, mk = r.mf_decode(g["stack"], 40)
    assert (mk == g["mask"]).all() and (has == g["has"]).all()
    assert (bits(ph[mk == 1]) == bits(g["phase"][mk == 1])).all()
    sc = golden("mf_scene")
    H, W = sc["stack"].shape[2:]
    cams, Q = cases.rig(W, H)
    pts, cnt = r.mf_triangulate(sc["phase"][0], sc["mask"][0], sc["phase"][1], sc["mask"][1], cams, Q, cases.RIGID)
    t = golden("mf_tri_rigid")
    assert (cnt == t["count"]).all() and (bits(pts[cnt > 0]) == bits(t["points"][cnt > 0])).all()
    assert (r.generate_mf(1280, 2) == golden("patterns")["mf_1280x2"]).all()
    # inputs are what cases.py generates today (numpy determinism)
    assert (cases.mf_pairs_stack() == g["stack"]).all() and (cases.mf_scene() == sc["stack"]).all()
