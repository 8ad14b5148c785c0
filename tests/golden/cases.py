"""Seeded inputs shared by make_ref_fixtures.py (runs the reference's own code on them) and test_ref_pin.py
(runs the oracle on them).  Small on purpose: the fixtures are committed."""
import numpy as np

import slr_b200
from slr_b200 import synth

RIGID = np.array([[0.98, -0.17, 0.05, 12.5], [0.17, 0.98, 0.02, -3.25], [-0.05, -0.01, 0.99, 40.0]], np.float32)


def mf_pairs_stack():
    """every a = G4-G2 in [-255,255] against 41 values of b = G1-G3 at frequency 0; shuffled pairs elsewhere"""
    bs = np.array([-255, -200, -128, -77, -64, -33, -17, -9, -5, -4, -3, -2, -1, 0, 1, 2, 3, 4, 5, 9, 17, 33, 64, 77,
                   128, 200, 255, -254, 254, -100, 100, -50, 50, -25, 25, -12, 12, -7, 7, -150, 150])
    a, b = np.meshgrid(np.arange(-255, 256), bs, indexing="ij")
    a, b = a.ravel(), b.ravel()
    n = a.size
    W = 512
    H = (n + W - 1) // W
    rng = np.random.default_rng(0)
    st = np.zeros((14, H * W), np.uint8)
    st[0] = 255
    perm = [np.arange(n), rng.permutation(n), rng.permutation(n)]
    for f in range(3):
        aa, bb = a[perm[f]], b[perm[f]]
        G = np.zeros((4, n), np.int64)
        G[3], G[1] = np.maximum(aa, 0), np.maximum(-aa, 0)
        G[0], G[2] = np.maximum(bb, 0), np.maximum(-bb, 0)
        st[2 + 4 * f:6 + 4 * f, :n] = G
        st[2 + 4 * f:6 + 4 * f, n:] = rng.integers(0, 256, (4, H * W - n))
    return st.reshape(14, H, W)


def mf_scene(W=192, H=10, seed=21, noise=2.0, intd=True):
    return synth.synth_mf(W, H, seed=seed, noise_dn=noise, integer_disparity=intd)


def gray_scene(W=160, H=12, seed=31, noise=3.0, rows=False, intd=False):
    return synth.synth_gray(W, H, seed=seed, noise_dn=noise, rows=rows, integer_disparity=intd)


def rig(W, H, distort=True):
    return slr_b200.synthetic_rig(W, H, distort=distort)


def gray_only_rig(W, H):
    """two converging cameras for the ray-ray path (rotation about y, baseline along x)"""
    c, s = np.cos(0.25), np.sin(0.25)
    camL = slr_b200.Camera(fc=(300.0, 300.0), cc=(W / 2, H / 2), dist=(-0.05, 0.01, 0.0005, -0.0003, 0.0),
                           R=(c, 0, s, 0, 1, 0, -s, 0, c), t=(-60.0, 0.0, 10.0))
    camR = slr_b200.Camera(fc=(305.0, 303.0), cc=(W / 2 + 1, H / 2 - 1), dist=(-0.04, 0.02, -0.0002, 0.0004, 0.0),
                           R=(c, 0, -s, 0, 1, 0, s, 0, c), t=(60.0, 0.0, 10.0))
    return [camL, camR]


def helper_points(n=64, seed=5):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 1280, (n, 2)).astype(np.float32), rng.normal(0, 1, (n, 4, 3)).astype(np.float32)


def mesh_cloud(w=29, h=17, seed=51, color=True):
    """A PointCloudImage with holes, multi-hit pixels, awkward magnitudes: sums [h,w,3], counts [h,w], colour u8."""
    rng = np.random.default_rng(seed)
    count = ((rng.random((h, w)) > 0.3) * rng.integers(1, 4, (h, w))).astype(np.uint8)
    count[0, 0] = 1                                # PLY vertex 0 exists: faces treat it as absent (meshcreator.cpp:79,100)
    pts = (rng.normal(0, 300, (h, w, 3)) * count[..., None]).astype(np.float32)
    pts[0, 0] = [1e-7, 123456789.0, -0.000123456]
    pts[1, 1] = [0.0, -0.0, 1e10]
    col = rng.integers(0, 256, (h, w, 3)).astype(np.uint8) if color else None
    return pts, count, col
