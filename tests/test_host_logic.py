"""CPU tests of the product's host logic: the C-ABI library loads, exports every symbol the header
declares, its host-side pattern generators agree with the oracle, and compute entry points fail
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import slr_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = slr_b200.capi()
    header = open(os.path.join(ROOT, "include", "slr_b200.h")).read()
    declared = set(re.findall(r"\b(slr_[a-z0-9_]+)\s*\(", header))
    declared -= {"slr_status", "slr_mode"}
    assert declared, "no declarations parsed"
    assert declared == set(slr_b200.EXPORTS), declared ^ set(slr_b200.EXPORTS)
    for name in sorted(declared):
        assert getattr(lib, name) is not None, name


def test_version_and_error_strings():
    lib = slr_b200.capi()
    assert b"sm_100a" in lib.slr_version()
    assert isinstance(lib.slr_last_error(), bytes)


@pytest.mark.parametrize("W,H,epi", [(1280, 8, True), (640, 6, True), (80, 72, False), (1280, 1024, False)])
def test_gray_patterns_match_oracle(oracle, W, H, epi):
    a = slr_b200.generate_gray_patterns(W, H, epi)
    b = oracle.generate_gray(W, H, epi)
    assert a.shape == b.shape and (a == b).all()


@pytest.mark.parametrize("W,H", [(1280, 4), (1024, 3), (912, 2)])
def test_mf_patterns_match_oracle(oracle, W, H):
    a = slr_b200.generate_mf_patterns(W, H)
    b = oracle.generate_mf(W, H)
    assert (a == b).all()


def test_gray_num_bits_matches_oracle(oracle):
    for n in (2, 3, 480, 640, 800, 1024, 1280, 2048, 3000, 4096):
        assert slr_b200.gray_num_bits(n) == oracle.gray_num_bits(n)


def test_strict_lookup_tables_equal_the_oracle_for_every_difference_pair(oracle):
    """The kernels decode strict mode through two tables (reciprocal multiplier per b, wrapped phase per sign and quotient)
    instead of the branches of Duke/mfreconstruct.cpp:246-261: every (G4-G2, G1-G3) pair against the C oracle and the
    numpy restatement, in the kernels' 2^-24 fixed point."""
    import ref_np
    fx = slr_b200.strict_tables()
    b, a = np.meshgrid(np.arange(-255, 256), np.arange(-255, 256), indexing="ij")
    P, ok = ref_np.wrapped_phase(a, b)
    want = np.where(ok, np.round(P.astype(np.float64) * 2.0 ** 24), -2.0 ** 31).astype(np.int64)
    assert (P.astype(np.float64) * 2.0 ** 24 == np.round(P.astype(np.float64) * 2.0 ** 24)).all()   # exact in fixed point
    assert (fx == want).all(), np.argwhere(fx != want)[:5]
    for bb in (-255, -128, -3, -1, 0, 1, 2, 7, 254, 255):          # the C oracle on a band of b values
        for aa in range(-255, 256):
            ok_o, p_o = oracle.wrapped_phase_strict(aa, bb)
            got = int(fx[bb + 255, aa + 255])
            assert (got == -2 ** 31) == (not ok_o), (aa, bb)
            if ok_o:
                assert got == int(round(float(np.float32(p_o)) * 2.0 ** 24)), (aa, bb)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = slr_b200.capi()
    h = C.c_void_p()
    st = lib.slr_create(C.byref(h), 0, 64, 16, 1)
    assert st != 0 and not h.value
    assert b"no CPU fallback" in lib.slr_last_error()
    with pytest.raises(slr_b200.SlrError):
        slr_b200.Engine(64, 16)


def test_bad_arguments_are_rejected():
    lib = slr_b200.capi()
    assert lib.slr_create(None, 0, 64, 16, 1) != 0
    h = C.c_void_p()
    assert lib.slr_create(C.byref(h), 0, 0, 16, 1) != 0
    assert lib.slr_generate_mf_patterns(None, 16, 16) != 0
    assert lib.slr_mf_decode(None, None, 1, 3, 4, 40, 0, None, None) != 0


def test_numpy_synth_shapes_and_determinism():
    from slr_b200 import synth
    a = synth.synth_mf(64, 16, seed=3)
    b = synth.synth_mf(64, 16, seed=3)
    assert a.shape == (2, 14, 16, 64) and (a == b).all()
    g = synth.synth_gray(64, 16, seed=3)
    assert g.shape == (2, 2 + 2 * 6, 16, 64)
    g2 = synth.synth_gray(64, 16, seed=3, rows=True)
    assert g2.shape == (2, 2 + 2 * 6 + 2 * 4, 16, 64)


def test_oracle_pipeline_on_numpy_synth(oracle):
    """Host-only end-to-end sanity of the test inputs: integer-disparity scenes give a dense cloud."""
    from slr_b200 import synth
    W, H = 256, 8
    st = synth.synth_mf(W, H, seed=5, integer_disparity=True)
    cams, Q = slr_b200.synthetic_rig(W, H)
    xyz, valid, mk, n = oracle.run_mf(st, cams, Q)
    assert n == valid.sum() and n > 0.5 * W * H
    assert np.isfinite(xyz[valid == 1]).all() and np.isnan(xyz[valid == 0]).all()


# ------------------------------------------------------------------------------------------------
# DotMatch::calMatrix's arithmetic (Duke/dotmatch.cpp:1180-1324) — host functions, no GPU needed
# ------------------------------------------------------------------------------------------------
def _kabsch(base, moving):
    """independent solution: rotation + translation moving -> base by SVD (Kabsch / Umeyama with unit scale)"""
    cb, cm = base.mean(0), moving.mean(0)
    Hm = (moving - cm).T @ (base - cb)
    U, _, Vt = np.linalg.svd(Hm)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1, 1, d]) @ U.T
    return R, cb - R @ cm


def _rot(axis, deg):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    a = np.deg2rad(deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


@pytest.mark.parametrize("n,noise,deg", [(3, 0.0, 20), (12, 0.0, 175), (40, 0.05, 63), (6, 0.0, 0.0)])
def test_horn_method_and_transfer_matrix_chain(n, noise, deg):
    rng = np.random.default_rng(n)
    R0, t0 = _rot([0.3, -1, 0.5], deg), np.array([12.5, -40.0, 310.0])
    moving = rng.normal(0, 80, (n, 3)) + [0, 0, 600]
    base = moving @ R0.T + t0 + rng.normal(0, noise, (n, 3))
    q7, scale = slr_b200.horn_method(base, moving, force_unit_scale=True)
    R, t = _kabsch(base, moving)
    w, x, y, z = q7[3:]
    assert w >= 0 and abs(w * w + x * x + y * y + z * z - 1) < 1e-12
    Rq = np.array([[w*w + x*x - y*y - z*z, 2*(x*y - w*z), 2*(x*z + w*y)],
                   [2*(x*y + w*z), w*w - x*x + y*y - z*z, 2*(y*z - w*x)],
                   [2*(z*x - w*y), 2*(z*y + w*x), w*w - x*x - y*y + z*z]])
    assert np.allclose(Rq, R, atol=1e-9) and np.allclose(q7[:3], t, atol=1e-6)
    assert abs(scale - 1) < (1e-9 if noise == 0 else 0.01)
    if noise == 0:
        assert np.allclose(Rq, R0, atol=1e-9) and np.allclose(q7[:3], t0, atol=1e-7)
    # MRPT's default: the translation carries the estimated scale  t = c_base - s R c_moving
    q7s, s = slr_b200.horn_method(base, moving)
    assert np.allclose(q7s[3:], q7[3:]) and np.allclose(q7s[:3], base.mean(0) - s * Rq @ moving.mean(0), atol=1e-9)
    # calMatrix: scan 1 writes the matrix itself, scan n > 1 chains it onto the accumulated one (:1299-1317)
    M1 = slr_b200.register_scan(base, moving)
    assert np.allclose(M1[:, :3], Rq, atol=1e-12) and np.allclose(M1[:, 3], q7s[:3], atol=1e-12)
    R2, t2 = _rot([1, 0.2, 0], 33), np.array([-5.0, 7.0, 90.0])
    moving2 = rng.normal(0, 60, (n, 3))
    base2 = moving2 @ R2.T + t2
    M2 = slr_b200.register_scan(base2, moving2, prev=M1)
    cur = slr_b200.register_scan(base2, moving2)
    assert np.allclose(M2[:, :3], M1[:, :3] @ cur[:, :3], atol=1e-12)
    assert np.allclose(M2[:, 3], M1[:, :3] @ cur[:, 3] + M1[:, 3], atol=1e-9)
    # a point of scan 2 lands where chaining the two motions puts it
    p = np.array([3.0, -2.0, 50.0])
    assert np.allclose(M2[:, :3] @ p + M2[:, 3], M1[:, :3] @ (cur[:, :3] @ p + cur[:, 3]) + M1[:, 3], atol=1e-9)


def test_horn_method_rejects_too_few_points():
    with pytest.raises(slr_b200.SlrError):
        slr_b200.horn_method(np.zeros((2, 3)), np.zeros((2, 3)))
