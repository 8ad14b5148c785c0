#!/usr/bin/env python
"""Runs the REFERENCE's own sources (oracle/_ref/libref.so = Duke/*.cpp compiled unmodified against
oracle/ref_shim/) on the seeded cases of cases.py and stores inputs + outputs as tests/golden/ref_*.npz.
Run in the build container (needs /root/reference):  make -C oracle ref && python tests/golden/make_ref_fixtures.py
The fixtures travel to the GPU box, where /root/reference does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import ref_lib  # noqa: E402


def main():
    r = ref_lib.load()
    out = {}

    # patterns
    out["patterns"] = dict(gray_epi_80x4=r.generate_gray(80, 4, True), gray_full_48x40=r.generate_gray(48, 40, False),
                           mf_1280x2=r.generate_mf(1280, 2),
                           layouts=np.array([r.gray_layout(w, h, e) for (w, h, e) in
                                             [(1280, 1024, 1), (1280, 1024, 0), (640, 480, 1), (2048, 1536, 1), (800, 600, 0)]]),
                           g2d=np.array([r.gray_to_dec([(g >> (10 - c)) & 1 for c in range(11)]) for g in range(0, 2048)]))

    # MF decode
    st = cases.mf_pairs_stack()
    ph, has, mk = r.mf_decode(st, 40)
    out["mf_pairs"] = dict(stack=st, phase=ph, has=has, mask=mk)
    sc = cases.mf_scene()
    dec = [r.mf_decode(sc[c], 40) for c in range(2)]
    out["mf_scene"] = dict(stack=sc, phase=np.stack([d[0] for d in dec]), has=np.stack([d[1] for d in dec]),
                           mask=np.stack([d[2] for d in dec]))

    # MF triangulate (distorted rig, with and without the rigid transform)
    H, W = sc.shape[2:]
    cams, Q = cases.rig(W, H)
    phs, hs = out["mf_scene"]["phase"], out["mf_scene"]["mask"]      # degenerate pixels excluded via final mask
    for name, rg in (("mf_tri", None), ("mf_tri_rigid", cases.RIGID)):
        pts, cnt = r.mf_triangulate(phs[0], hs[0], phs[1], hs[1], cams, Q, rg)
        out[name] = dict(points=pts, count=cnt)
    pts, cnt = r.mf_triangulate(phs[0], hs[0], phs[1], hs[1], cams, Q, None, scan=(6, 100))   # F7 drop rule
    out["mf_tri_drop"] = dict(points=pts, count=cnt)

    # Gray EPI decode + GE triangulate
    gs = cases.gray_scene()
    Hg, Wg = gs.shape[2:]
    nc = r.gray_layout(Wg, Hg, True)[1]
    gd = [r.gray_decode(gs[c], nc, 0, 40, 5, Wg, Hg) for c in range(2)]
    out["ge_decode"] = dict(stack=gs, col=np.stack([d[0] for d in gd]), mask=np.stack([d[2] for d in gd]))
    _, Qg = cases.rig(Wg, Hg)
    cols, mks = out["ge_decode"]["col"], out["ge_decode"]["mask"]
    pts, cnt, color = r.ge_triangulate(cols[0], mks[0], cols[1], mks[1], Qg, cases.RIGID, gs[0, 0], gs[1, 0])
    out["ge_tri"] = dict(points=pts, count=cnt, color=color)

    # Gray-only decode + bucket triangulation
    go = cases.gray_scene(W=48, H=40, seed=41, noise=2.0, rows=True, intd=True)
    Ho, Wo = go.shape[2:]
    _, nco, nro = r.gray_layout(Wo, Ho, False)
    od = [r.gray_decode(go[c], nco, nro, 40, 3, Wo, Ho) for c in range(2)]
    out["go_decode"] = dict(stack=go, col=np.stack([d[0] for d in od]), row=np.stack([d[1] for d in od]),
                            mask=np.stack([d[2] for d in od]))
    camsg = cases.gray_only_rig(Wo, Ho)
    d = out["go_decode"]
    has = (d["col"] >= 0)
    ssum, cnt = r.gray_triangulate(d["col"][0], d["row"][0], has[0], d["col"][1], d["row"][1], has[1], Wo, Ho, camsg, None)
    out["go_tri"] = dict(sum=ssum, cnt=cnt)

    # helpers
    px, vecs = cases.helper_points()
    cams, _ = cases.rig(1280, 1024)
    und = np.array([[r.undistort(x, y, cams[c]) for (x, y) in px] for c in range(2)], np.float32)
    ll = [r.line_line(v[0] * 50, v[1] / np.linalg.norm(v[1]), v[2] * 50, v[3] / np.linalg.norm(v[3])) for v in vecs]
    c2w = np.array([r.cam2world(camsg[i % 2], v[0] * 10) for i, v in enumerate(vecs)], np.float32)
    nrm = np.array([r.normalize(v[1] * (1e-7 if i % 9 == 0 else 3.0)) for i, v in enumerate(vecs)], np.float32)
    rng = np.random.default_rng(9)
    iw, jh = rng.integers(0, 7, 600).astype(np.int32), rng.integers(0, 5, 600).astype(np.int32)
    iw[:300], jh[:300] = 2, 3                    # > 255 additions into one cell: the u8 count wraps (pointcloudimage.cpp:95)
    pp = rng.normal(0, 10, (600, 3)).astype(np.float32)
    pcs, pcc = r.pointcloud_add(6, 4, iw, jh, pp)
    out["helpers"] = dict(px=px, vecs=vecs, undistort=und, ll_ok=np.array([o for o, _ in ll]), ll_p=np.array([p for _, p in ll]),
                          cam2world=c2w, normalize=nrm, pc_iw=iw, pc_jh=jh, pc_pts=pp, pc_sum=pcs, pc_cnt=pcc)

    # N3: MeshCreator text export
    import tempfile
    mesh = {}
    for tag, color in (("c", True), ("n", False)):
        pts, cnt, col = cases.mesh_cloud(color=color)
        h, w = cnt.shape
        for obj in (False, True):
            with tempfile.NamedTemporaryFile(suffix=".txt") as f:
                r.export_mesh(pts, cnt, w, h, f.name, obj, col)
                mesh[f"{tag}_{'obj' if obj else 'ply'}"] = np.frombuffer(open(f.name, "rb").read(), np.uint8)
        for off in (True, False):
            with tempfile.NamedTemporaryFile(suffix=".txt") as f:
                r.export_xyz(pts, cnt, w, h, f.name, off, True, col)
                mesh[f"{tag}_xyz_{'off' if off else 'on'}"] = np.frombuffer(open(f.name, "rb").read(), np.uint8)
    out["mesh"] = mesh

    for name, arrs in out.items():
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **arrs)
        print(f"ref_{name}.npz", {k: v.shape for k, v in arrs.items()})


if __name__ == "__main__":
    main()
