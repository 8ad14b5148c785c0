"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the
same inputs.  Bars: Gray indices / masks / match columns bit-exact; strict-mode phase and XYZ are
also compared bit-for-bit (the kernels use the same IEEE operation sequence), with the north-star
tolerance (1e-4 relative) written as the fallback bar for floating point; corrected mode 1e-4."""
import numpy as np
import pytest

import slr_b200
from slr_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-4          # BASELINE.json north_star: "within 1e-4 relative on unwrapped phase and XYZ"


def _t(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_float_parity(gpu, cpu, what):
    gpu, cpu = np.asarray(gpu), np.asarray(cpu)
    assert (np.isnan(gpu) == np.isnan(cpu)).all(), f"{what}: NaN pattern differs"
    ok = ~np.isnan(cpu)
    assert np.allclose(gpu[ok], cpu[ok], rtol=RTOL, atol=1e-6), f"{what}: outside 1e-4 relative"
    exact = (bits(gpu) == bits(cpu)).mean()
    assert exact == 1.0, f"{what}: only {exact:.6f} of values bit-exact"


def all_pairs_stack():
    """A 14-plane stack whose pixels realise every (a, b) = (G4-G2, G1-G3) in [-255, 255]^2 at frequency 0
    and shuffled pairs at frequencies 1, 2 (exhaustive for getPhase's per-frequency branch table)."""
    a, b = np.meshgrid(np.arange(-255, 256), np.arange(-255, 256), indexing="ij")
    a, b = a.ravel(), b.ravel()
    n = a.size                                   # 261121
    W = 512
    H = (n + W - 1) // W
    pad = W * H - n
    rng = np.random.default_rng(0)
    st = np.zeros((14, H * W), np.uint8)
    st[0] = 255
    st[1] = 0
    perm = [np.arange(n), rng.permutation(n), rng.permutation(n)]
    for f in range(3):
        aa, bb = a[perm[f]], b[perm[f]]
        G = np.zeros((4, n), np.int64)
        G[3] = np.maximum(aa, 0)                 # G4
        G[1] = np.maximum(-aa, 0)                # G2
        G[0] = np.maximum(bb, 0)                 # G1
        G[2] = np.maximum(-bb, 0)                # G3
        st[2 + 4 * f:6 + 4 * f, :n] = G
        st[2 + 4 * f:6 + 4 * f, n:] = rng.integers(0, 256, (4, pad))
    return st.reshape(14, H, W), W, H


def test_k1_strict_all_difference_pairs(cuda_engine_factory, oracle):
    st, W, H = all_pairs_stack()
    eng = cuda_engine_factory(W, H)
    stack = np.stack([st, st[:, ::-1].copy()])[None]            # [1, 2, 14, H, W]
    ph, mk = eng.mf_decode(_t(stack), black_thr=40)
    for cam in range(2):
        ph_o, mk_o = oracle.mf_decode(stack[0, cam], black_thr=40)
        assert (mk[0, cam].cpu().numpy() == mk_o).all()
        assert_float_parity(ph[0, cam].cpu().numpy(), ph_o, f"strict phase cam{cam}")


@pytest.mark.parametrize("W,H,noise", [(1280, 64, 0.0), (1280, 48, 2.0), (640, 480, 1.0), (20, 3, 0.0)])
def test_k1_strict_synth_and_scalar_path(cuda_engine_factory, oracle, W, H, noise):
    eng = cuda_engine_factory(W, H, 2)
    stack = np.stack([synth.synth_mf(W, H, seed=s, noise_dn=noise) for s in (1, 2)])
    ph, mk = eng.mf_decode(_t(stack), black_thr=40)
    for b in range(2):
        for cam in range(2):
            ph_o, mk_o = oracle.mf_decode(stack[b, cam], black_thr=40)
            assert (mk[b, cam].cpu().numpy() == mk_o).all()
            assert_float_parity(ph[b, cam].cpu().numpy(), ph_o, "strict phase")


@pytest.mark.parametrize("F,S", [(3, 4), (4, 8), (3, 5)])
def test_k1_corrected_mode(cuda_engine_factory, oracle, F, S):
    W, H = 256, 32
    eng = cuda_engine_factory(W, H)
    rng = np.random.default_rng(F * 10 + S)
    # smooth fringes so the wrap points are sparse
    x = np.arange(W)[None, None, None, :] / W
    stack = np.zeros((1, 2, 2 + F * S, H, W), np.uint8)
    stack[:, :, 0] = 220
    stack[:, :, 1] = 10
    freqs = [70, 64, 59, 55][:F]
    for f in range(F):
        for s in range(S):
            v = 128 + 90 * np.cos(2 * np.pi * freqs[f] * (x + 0.013) + 2 * np.pi * s / S)
            stack[:, :, 2 + S * f + s] = np.clip(v + rng.normal(0, 1.0, (1, 2, H, W)), 0, 255).astype(np.uint8)
    ph, mk = eng.mf_decode(_t(stack), F=F, S=S, black_thr=40, mode=slr_b200.MODE_CORRECTED)
    for cam in range(2):
        ph_o, mk_o = oracle.mf_decode(stack[0, cam], F=F, S=S, black_thr=40, mode=1)
        g = ph[0, cam].cpu().numpy()
        assert (mk[0, cam].cpu().numpy() == mk_o).all()
        ok = mk_o == 1
        d = np.abs(g[ok] - ph_o[ok])
        d = np.minimum(d, 255.0 - d)                            # the phase is circular with period 255
        assert (d <= RTOL * 255.0).all(), d.max()


def decode_oracle(oracle, stack):
    """stack [B,2,14,H,W] -> phase [B,2,H,W], mask"""
    B = stack.shape[0]
    ph = np.empty(stack.shape[:2] + stack.shape[3:], np.float32)
    mk = np.empty(ph.shape, np.uint8)
    for b in range(B):
        for cam in range(2):
            ph[b, cam], mk[b, cam] = oracle.mf_decode(stack[b, cam], black_thr=40)
    return ph, mk


@pytest.mark.parametrize("W,H,intd,noise,rigid", [(1280, 32, True, 0.0, False), (1280, 24, False, 1.5, True),
                                                  (640, 40, True, 2.0, False), (64, 8, True, 0.0, True)])
def test_k3a_phase_match_vs_oracle(cuda_engine_factory, oracle, W, H, intd, noise, rigid):
    eng = cuda_engine_factory(W, H, 2)
    cams, Q = slr_b200.synthetic_rig(W, H)
    M = np.array([[0.98, -0.17, 0.05, 12.5], [0.17, 0.98, 0.02, -3.25], [-0.05, -0.01, 0.99, 40.0]], np.float32) \
        if rigid else None
    eng.set_calib(cams, Q, M)
    stack = np.stack([synth.synth_mf(W, H, seed=s, integer_disparity=intd, noise_dn=noise) for s in (11, 12)])
    ph, mk = decode_oracle(oracle, stack)
    xyz, valid, k, n = eng.match_triangulate_phase(_t(ph), _t(mk))
    total = 0
    for b in range(2):
        xyz_o, valid_o, k_o, n_o = oracle.mf_triangulate(ph[b, 0], mk[b, 0], ph[b, 1], mk[b, 1], cams, Q, M)
        assert (k[b].cpu().numpy() == k_o).all(), "match column differs"
        assert (valid[b].cpu().numpy() == valid_o).all()
        assert_float_parity(xyz[b].cpu().numpy(), xyz_o, "XYZ")
        total += n_o
    assert int(n.item()) == total
    assert total > 0


def test_k3a_dense_q_and_zero_disparity(cuda_engine_factory, oracle):
    """A dense (non-stereoRectify-shaped) Q takes the generic emitter; a zero disparity with Q[3][3] == 0
    divides by zero exactly as the reference's double arithmetic does (inf / nan patterns must agree)."""
    W, H = 64, 4
    eng = cuda_engine_factory(W, H)
    cams, Q = slr_b200.synthetic_rig(W, H, distort=False)
    cams = [slr_b200.Camera(fc=(900, 900), cc=(32, 2)), slr_b200.Camera(fc=(900, 900), cc=(32, 2))]
    rng = np.random.default_rng(5)
    ph = np.zeros((1, 2, H, W), np.float32)
    ph[0, :, :] = (np.arange(W, dtype=np.float32) * 0.7)[None, None, :]      # zero disparity everywhere
    ph[0, 0, 2:, 5:] = ph[0, 1, 2:, :-5]                                     # rows 2,3: disparity 5
    mk = np.ones((1, 2, H, W), np.uint8)
    for Qm in (Q, Q + rng.normal(0, 0.01, (4, 4)), np.where(Q == 0, -0.0, Q)):
        eng.set_calib(cams, Qm)
        xyz, valid, k, n = eng.match_triangulate_phase(_t(ph), _t(mk))
        xyz_o, valid_o, k_o, n_o = oracle.mf_triangulate(ph[0, 0], mk[0, 0], ph[0, 1], mk[0, 1], cams, Qm)
        assert (k[0].cpu().numpy() == k_o).all() and int(n.item()) == n_o
        assert (bits(xyz[0].cpu().numpy()) == bits(xyz_o)).all()


def test_k3a_edge_cases(cuda_engine_factory, oracle):
    W, H = 64, 6
    eng = cuda_engine_factory(W, H)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    rng = np.random.default_rng(3)
    ph = np.zeros((1, 2, H, W), np.float32)
    mk = np.ones((1, 2, H, W), np.uint8)
    ph[0, :, 0] = 7.25                                             # row 0: every pixel the same phase
    mk[0, :, 1] = 0                                                # row 1: nothing carries a phase
    ph[0, :, 2] = rng.choice([1.0, 1.05, 1.12, 1.2, 9.0], (2, W))  # row 2: heavy duplicates near the tolerance
    ph[0, :, 3] = rng.uniform(-200, 480, (2, W))                   # row 3: strict-mode value range
    ph[0, 0, 4] = np.nan                                           # row 4: NaN phase under a set mask
    ph[0, 1, 4] = rng.uniform(0, 1, W)
    ph[0, :, 5] = rng.uniform(0, 0.5, (2, W))                      # row 5: everything within tolerance chains
    mk[0, 1, 5, ::3] = 0
    xyz, valid, k, n = eng.match_triangulate_phase(_t(ph), _t(mk))
    xyz_o, valid_o, k_o, n_o = oracle.mf_triangulate(ph[0, 0], mk[0, 0], ph[0, 1], mk[0, 1], cams, Q)
    assert (k[0].cpu().numpy() == k_o).all()
    assert (valid[0].cpu().numpy() == valid_o).all()
    assert_float_parity(xyz[0].cpu().numpy(), xyz_o, "XYZ edge")
    assert int(n.item()) == n_o


def test_run_mf_full_frame_vs_oracle(cuda_engine_factory, oracle):
    """BASELINE config 3's MF half at full size: 1280x1024, one scan, fused entry point."""
    W, H = 1280, 1024
    eng = cuda_engine_factory(W, H, 2)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = eng.synth_mf(2, seed=42, integer_disparity=True, noise_dn=0.0)
    xyz, valid, k, n = eng.run_mf(stack, black_thr=40)
    hs = stack.cpu().numpy()
    tot = 0
    for b in range(2):
        xyz_o, valid_o, k_o, n_o = oracle.run_mf(hs[b], cams, Q, nthreads=oracle.max_threads())
        assert (k[b].cpu().numpy() == k_o).all()
        assert (valid[b].cpu().numpy() == valid_o).all()
        assert_float_parity(xyz[b].cpu().numpy(), xyz_o, "XYZ full frame")
        tot += n_o
    assert int(n.item()) == tot and tot > 0.5 * 2 * W * H
    # size-independent properties: determinism / idempotence, count == number of valid flags
    xyz2, valid2, k2, n2 = eng.run_mf(stack, black_thr=40)
    assert (bits(xyz2.cpu().numpy()) == bits(xyz.cpu().numpy())).all() and int(n2.item()) == int(valid2.sum().item())


def test_run_mf_host_matches_device(cuda_engine_factory):
    W, H = 640, 96
    eng = cuda_engine_factory(W, H, 2)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = eng.synth_mf(3, seed=7, noise_dn=1.0)
    xyz, valid, k, n = eng.run_mf(stack)
    h_stack = stack.cpu().numpy()
    h_xyz = np.empty((3, H, W, 3), np.float32)
    h_valid = np.empty((3, H, W), np.uint8)
    h_k = np.empty((3, H, W), np.int32)
    n_host = eng.run_mf_host(h_stack, h_xyz, h_valid, h_k)
    assert n_host == int(n.item())
    assert (bits(h_xyz) == bits(xyz.cpu().numpy())).all()
    assert (h_valid == valid.cpu().numpy()).all() and (h_k == k.cpu().numpy()).all()


# ---------------------------------------------------------------------------------------------
# Gray path
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H,rows,white_thr,noise", [(640, 480, False, 0, 0.0), (1280, 32, False, 10, 3.0),
                                                      (96, 80, True, 5, 2.0), (20, 3, True, 0, 0.0)])
def test_k2_gray_decode_vs_oracle(cuda_engine_factory, oracle, W, H, rows, white_thr, noise):
    """Config 1 (640x480, 10 bitplanes) and friends: indices and masks bit-exact."""
    eng = cuda_engine_factory(W, H)
    stack = synth.synth_gray(W, H, seed=5, noise_dn=noise, rows=rows)[None]
    nc = oracle.gray_num_bits(W)
    nr = oracle.gray_num_bits(H) if rows else 0
    col, row, mk = eng.gray_decode(_t(stack), nc, nr, black_thr=40, white_thr=white_thr, scan_w=W, scan_h=H)
    for cam in range(2):
        c_o, r_o, m_o = oracle.gray_decode(stack[0, cam], nc, nr, 40, white_thr, W, H)
        assert (col[0, cam].cpu().numpy() == c_o).all()
        assert (mk[0, cam].cpu().numpy() == m_o).all()
        if rows:
            assert (row[0, cam].cpu().numpy() == r_o).all()


@pytest.mark.parametrize("W,H,intd,noise,color", [(1280, 32, True, 0.0, False), (640, 48, False, 4.0, True),
                                                  (64, 8, True, 6.0, True)])
def test_k3b_code_match_vs_oracle(cuda_engine_factory, oracle, W, H, intd, noise, color):
    eng = cuda_engine_factory(W, H)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = synth.synth_gray(W, H, seed=9, integer_disparity=intd, noise_dn=noise)[None]
    nc = oracle.gray_num_bits(W)
    cols, mks = [], []
    for cam in range(2):
        c_o, _, m_o = oracle.gray_decode(stack[0, cam], nc, 0, 40, 0, W, H)
        cols.append(c_o)
        mks.append(m_o)
    col = np.stack(cols)[None]
    mk = np.stack(mks)[None]
    white = np.ascontiguousarray(stack[:, :, 0]) if color else None
    xyz, valid, k, colr, n = eng.match_triangulate_code(_t(col), _t(mk), _t(white) if color else None)
    xyz_o, valid_o, k_o, col_o, n_o = oracle.ge_triangulate(col[0, 0], mk[0, 0], col[0, 1], mk[0, 1], Q,
                                                            whiteL=white[0, 0] if color else None,
                                                            whiteR=white[0, 1] if color else None)
    assert (k[0].cpu().numpy() == k_o).all()
    assert (valid[0].cpu().numpy() == valid_o).all()
    assert_float_parity(xyz[0].cpu().numpy(), xyz_o, "GE XYZ")
    if color:
        assert (colr[0].cpu().numpy() == col_o).all()
    assert int(n.item()) == n_o and n_o > 0


def test_k3b_adversarial_chains(cuda_engine_factory, oracle):
    """Rows built to stress the kstart chain: sparse matches, repeated codes, decreasing codes."""
    W, H = 256, 8
    eng = cuda_engine_factory(W, H)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    rng = np.random.default_rng(1)
    col = np.full((1, 2, H, W), -1, np.int32)
    col[0, :, 0] = rng.integers(0, 4, (2, W))                       # few codes, many repeats
    col[0, 0, 1] = np.arange(W)[::-1]                               # decreasing left codes
    col[0, 1, 1] = np.arange(W)
    col[0, 0, 2, 200] = 7                                           # a lone left pixel at the far end
    col[0, 1, 2] = 7
    col[0, :, 3] = rng.integers(0, W, (2, W))                       # random
    col[0, 0, 4] = np.repeat(np.arange(W // 4), 4)                  # runs
    col[0, 1, 4] = np.repeat(np.arange(W // 2), 2)
    col[0, 0, 5, ::17] = 3                                          # sparse left, all-matching right
    col[0, 1, 5] = 3
    col[0, 0, 6] = 5                                                # left all same, right has a single 5 late
    col[0, 1, 6, 250] = 5
    mk = (col >= 0).astype(np.uint8)
    xyz, valid, k, _, n = eng.match_triangulate_code(_t(col), _t(mk))
    xyz_o, valid_o, k_o, _, n_o = oracle.ge_triangulate(col[0, 0], mk[0, 0], col[0, 1], mk[0, 1], Q)
    assert (k[0].cpu().numpy() == k_o).all()
    assert (valid[0].cpu().numpy() == valid_o).all()
    assert_float_parity(xyz[0].cpu().numpy(), xyz_o, "GE XYZ adversarial")
    assert int(n.item()) == n_o


def test_run_ge_full_frame_vs_oracle(cuda_engine_factory, oracle):
    W, H = 1280, 1024
    eng = cuda_engine_factory(W, H)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = eng.synth_gray(1, seed=4, integer_disparity=False, noise_dn=2.0)
    nc = oracle.gray_num_bits(W)
    xyz, valid, k, colr, n = eng.run_ge(stack, nc, black_thr=40, white_thr=3, have_color=True)
    hs = stack.cpu().numpy()
    cols, mks = zip(*[(lambda r: (r[0], r[2]))(oracle.gray_decode(hs[0, cam], nc, 0, 40, 3, W, H)) for cam in range(2)])
    xyz_o, valid_o, k_o, col_o, n_o = oracle.ge_triangulate(cols[0], mks[0], cols[1], mks[1], Q,
                                                            whiteL=hs[0, 0, 0], whiteR=hs[0, 1, 0],
                                                            nthreads=oracle.max_threads())
    assert (k[0].cpu().numpy() == k_o).all()
    assert (valid[0].cpu().numpy() == valid_o).all()
    assert (colr[0].cpu().numpy() == col_o).all()
    assert_float_parity(xyz[0].cpu().numpy(), xyz_o, "GE XYZ full frame")
    assert int(n.item()) == n_o and n_o > 0.3 * W * H


@pytest.mark.parametrize("mode,F,S,W,H", [(0, 3, 4, 1280, 40), (1, 3, 4, 1280, 40), (1, 4, 8, 640, 24), (0, 3, 4, 2048, 16),
                                         (0, 3, 4, 4096, 8), (0, 3, 4, 48, 5)])
def test_fused_equals_unfused_kernels(cuda_engine_factory, mode, F, S, W, H):
    """slr_run_mf (one fused kernel where the shape allows) against K1 + K3a run back to back: same device
    arithmetic, so every output must be identical, in both decode modes and for widths that take the
    two-chunk (2048), un-fused fallback (4096) and tiny-row (48) routes."""
    import torch
    eng = cuda_engine_factory(W, H, 2)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    rng = np.random.default_rng(W + F)
    if F == 3 and S == 4:
        stack = np.stack([synth.synth_mf(W, H, seed=s, noise_dn=1.0, integer_disparity=(mode == 0)) for s in (3, 4)])
    else:
        stack = rng.integers(0, 256, (2, 2, 2 + F * S, H, W)).astype(np.uint8)
        stack[:, :, 0] = 230
        stack[:, :, 1] = 12
    st = _t(stack)
    xyz, valid, k, n = eng.run_mf(st, F=F, S=S, black_thr=40, mode=mode)
    ph, mk = eng.mf_decode(st, F=F, S=S, black_thr=40, mode=mode)
    xyz2, valid2, k2, n2 = eng.match_triangulate_phase(ph, mk)
    assert torch.equal(k, k2) and torch.equal(valid, valid2)
    assert (bits(xyz.cpu().numpy()) == bits(xyz2.cpu().numpy())).all()
    assert int(n.item()) == int(n2.item()) == int(valid.sum().item())


# ---------------------------------------------------------------------------------------------
# Gray-only (un-rectified, projector-cell buckets, ray-ray midpoints)
# ---------------------------------------------------------------------------------------------
def _cases():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import cases
    return cases


@pytest.mark.parametrize("W,H,rigid,noise", [(48, 40, False, 2.0), (96, 64, True, 3.0)])
def test_k3c_bucket_triangulate_vs_oracle(cuda_engine_factory, oracle, W, H, rigid, noise):
    cases = _cases()
    eng = cuda_engine_factory(W, H, 2)
    cams = cases.gray_only_rig(W, H)
    _, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q, cases.RIGID if rigid else None)
    stack = np.stack([synth.synth_gray(W, H, seed=s, noise_dn=noise, rows=True, integer_disparity=True) for s in (41, 42)])
    nc, nr = oracle.gray_num_bits(W), oracle.gray_num_bits(H)
    col, row, mk = eng.gray_decode(_t(stack), nc, nr, black_thr=40, white_thr=3, scan_w=W, scan_h=H)
    ssum, cnt, n = eng.bucket_triangulate(col, row, mk, W, H)
    total = 0
    for b in range(2):
        d = [oracle.gray_decode(stack[b, cam], nc, nr, 40, 3, W, H) for cam in range(2)]
        for cam in range(2):
            assert (col[b, cam].cpu().numpy() == d[cam][0]).all() and (row[b, cam].cpu().numpy() == d[cam][1]).all()
        s_o, c_o, n_o = oracle.gray_triangulate(d[0][0], d[0][1], d[0][2], d[1][0], d[1][1], d[1][2], W, H, cams,
                                                cases.RIGID if rigid else None)
        assert (cnt[b].cpu().numpy() == c_o).all()
        g = ssum[b].cpu().numpy()
        assert (bits(g[c_o > 0]) == bits(s_o[c_o > 0])).all()
        total += n_o
    assert int(n.item()) == total and total > 20


def test_k3c_many_points_per_cell_wraps_like_the_reference(cuda_engine_factory, oracle):
    """All camera pixels decode to a handful of projector cells: > 255 pair midpoints per cell, so the reference's
    u8 count wraps and the sum restarts (pointcloudimage.cpp:91-95); ordering of the pair walk matters."""
    cases = _cases()
    W, H = 32, 24
    eng = cuda_engine_factory(W, H)
    cams = cases.gray_only_rig(W, H)
    _, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    rng = np.random.default_rng(2)
    col = rng.integers(0, 2, (1, 2, H, W)).astype(np.int32)          # 2 x 2 cells, ~190 pixels each per camera
    row = rng.integers(0, 2, (1, 2, H, W)).astype(np.int32)
    mk = (rng.random((1, 2, H, W)) < 0.9).astype(np.uint8)
    col[mk == 0] = -1
    row[mk == 0] = -1
    ssum, cnt, n = eng.bucket_triangulate(_t(col), _t(row), _t(mk), W, H)
    s_o, c_o, n_o = oracle.gray_triangulate(col[0, 0], row[0, 0], mk[0, 0], col[0, 1], row[0, 1], mk[0, 1], W, H, cams)
    assert (cnt[0].cpu().numpy() == c_o).all() and int(n.item()) == n_o
    assert (bits(ssum[0].cpu().numpy()[c_o > 0]) == bits(s_o[c_o > 0])).all()


# ---------------------------------------------------------------------------------------------
# K0: rectification on load — stereoRect::doStereoRectify == cv::remap(INTER_LINEAR)
# ---------------------------------------------------------------------------------------------
def _warp_maps(W, H, seed):
    cv2 = pytest.importorskip("cv2")
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    maps1, maps2 = [], []
    for cam in range(2):
        mx = xs * (1.01 - 0.02 * cam) - 5 + 3 * np.sin(ys / 17 + seed)
        my = ys * 0.99 + 2 - 4 * cam + 2 * np.cos(xs / 23)
        mx[:3], my[:3] = xs[:3], ys[:3]                  # exact integer coordinates: the (0,0) weight entry
        m1, m2 = cv2.convertMaps(mx, my, cv2.CV_16SC2)
        maps1.append(m1)
        maps2.append(m2)
    return np.stack(maps1), np.stack(maps2)


def test_k0_rectify_vs_oracle_and_opencv(cuda_engine_factory, oracle):
    cv2 = pytest.importorskip("cv2")
    W, H, N = 320, 96, 5
    eng = cuda_engine_factory(W, H, 2)
    m1, m2 = _warp_maps(W, H, 0.3)
    eng.set_rectify_maps(m1, m2)
    rng = np.random.default_rng(8)
    raw = rng.integers(0, 256, (2, 2, N, H, W)).astype(np.uint8)
    out = eng.rectify_stack(_t(raw)).cpu().numpy()
    for b in range(2):
        for cam in range(2):
            for n in range(N):
                exp = oracle.remap_linear(raw[b, cam, n], m1[cam], m2[cam])
                assert (out[b, cam, n] == exp).all()
            assert (cv2.remap(raw[b, cam, 0], m1[cam], m2[cam], cv2.INTER_LINEAR) == out[b, cam, 0]).all()


def test_host_pipeline_with_raw_input_equals_rectify_then_run(cuda_engine_factory):
    W, H = 640, 48
    eng = cuda_engine_factory(W, H, 2)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    m1, m2 = _warp_maps(W, H, 1.1)
    eng.set_rectify_maps(m1, m2)
    raw = np.stack([synth.synth_mf(W, H, seed=s, noise_dn=1.0) for s in (5, 6, 7)])
    rect = eng.rectify_stack(_t(raw))
    xyz, valid, k, n = eng.run_mf(rect)
    h_xyz = np.empty((3, H, W, 3), np.float32)
    h_valid = np.empty((3, H, W), np.uint8)
    h_k = np.empty((3, H, W), np.int32)
    eng.set_host_input_raw(True)
    n_host = eng.run_mf_host(raw, h_xyz, h_valid, h_k)
    eng.set_host_input_raw(False)
    assert n_host == int(n.item()) and n_host > 0
    assert (bits(h_xyz) == bits(xyz.cpu().numpy())).all() and (h_k == k.cpu().numpy()).all()


def test_wide_rows_take_the_plain_kernels_and_still_match_the_oracle(cuda_engine_factory, oracle):
    """4096-wide rows (BASELINE config 5's width) exceed the fused kernel's shared-memory stage: slr_run_mf runs
    K1 + the plain K3a kernel.  Same exactness bar."""
    W, H = 4096, 3
    eng = cuda_engine_factory(W, H)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = synth.synth_mf(W, H, seed=77, noise_dn=1.0)[None]
    xyz, valid, k, n = eng.run_mf(_t(stack))
    xyz_o, valid_o, k_o, n_o = oracle.run_mf(stack[0], cams, Q, nthreads=oracle.max_threads())
    assert (k[0].cpu().numpy() == k_o).all() and int(n.item()) == n_o and n_o > 1000
    assert (bits(xyz[0].cpu().numpy()) == bits(xyz_o)).all()


# ------------------------------------------------------------------------------------------------
# K5: mesh indexing (MeshCreator's vertex numbering + faces, Duke/meshcreator.cpp:16-166)
# ------------------------------------------------------------------------------------------------
def _random_cloud(w, h, seed, fill=0.7):
    rng = np.random.default_rng(seed)
    count = ((rng.random((h, w)) < fill) * rng.integers(1, 4, (h, w))).astype(np.uint8)
    pts = (rng.normal(0, 300, (h, w, 3)) * count[..., None]).astype(np.float32)
    return pts, count


@pytest.mark.parametrize("w,h,fill", [(29, 17, 0.7), (1, 1, 1.0), (5, 1, 1.0), (1, 7, 0.5), (64, 48, 0.0), (333, 257, 0.9),
                                      (2048, 3, 0.6)])
@pytest.mark.parametrize("first_vertex", [0, 1])
def test_k5_mesh_index_vs_oracle(cuda_engine_factory, oracle, w, h, fill, first_vertex):
    eng = cuda_engine_factory(64, 16)
    pts, count = _random_cloud(w, h, 7 * w + h, fill)
    if w * h > 1:
        count[0, 0] = 1   # PLY numbering: vertex 0 exists and must read as "absent" when faces are formed
    vert, src, faces = eng.mesh_index(_t(pts), _t(count), first_vertex)
    vo, so, fo = oracle.mesh_index(pts, count, w, h, first_vertex)
    assert vert.shape[0] == len(vo) and faces.shape[0] == len(fo)
    assert (bits(vert.cpu().numpy()) == bits(vo)).all()
    assert (src.cpu().numpy() == so).all()
    assert (faces.cpu().numpy() == fo).all()


def test_k5_mesh_index_full_frame_cloud(cuda_engine_factory, oracle):
    """The mesh of a real 1280x1024 MF cloud: PointCloudImage(1280, 1024) filled as MFReconstruct does (F7 drop rule)."""
    W, H = 1280, 1024
    eng = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = eng.synth_mf(1, seed=5, integer_disparity=True, noise_dn=0.0)
    xyz, valid, k, n = eng.run_mf(stack, black_thr=40)
    sums, count = oracle.pointcloud_from_dense(xyz[0].cpu().numpy(), valid[0].cpu().numpy(), W, H)
    for first_vertex in (0, 1):
        vert, src, faces = eng.mesh_index(_t(sums), _t(count), first_vertex)
        vo, so, fo = oracle.mesh_index(sums, count, W, H, first_vertex)
        assert len(vo) > 0.5 * 1024 * 1024 and len(fo) > len(vo)
        assert (bits(vert.cpu().numpy()) == bits(vo)).all() and (src.cpu().numpy() == so).all()
        assert (faces.cpu().numpy() == fo).all()
    # size-independent properties: vertex numbers are 0/1-based ranks, every face references existing vertices,
    # face count never exceeds two per pixel
    f = faces.cpu().numpy()
    assert f.min() >= 1 and f.max() <= len(vo) and len(f) <= 2 * W * H


# ------------------------------------------------------------------------------------------------
# widths that are not a multiple of 16 (the match kernels' TMA rows): zero-padded copies on a child engine
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H", [(1000, 12), (52, 9), (37, 7), (1290, 6)])
def test_ragged_widths_match_the_oracle(cuda_engine_factory, oracle, W, H):
    """The reference takes any image size; here rows move by TMA in 16-byte multiples, so other widths run padded.
    Same exactness bar for the MF and GE pipelines, the phase / code match entry points and the host entry point."""
    eng = cuda_engine_factory(W, H, 2)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    # ---- MF: device entry, un-fused entry points, host entry ----
    stack = np.stack([synth.synth_mf(W, H, seed=s, noise_dn=1.0) for s in (5, 6)])
    xyz, valid, k, n = eng.run_mf(_t(stack), black_thr=40)
    tot = 0
    for b in range(2):
        xyz_o, valid_o, k_o, n_o = oracle.run_mf(stack[b], cams, Q)
        assert (k[b].cpu().numpy() == k_o).all() and (valid[b].cpu().numpy() == valid_o).all()
        assert (bits(xyz[b].cpu().numpy()) == bits(xyz_o)).all()
        tot += n_o
    assert int(n.item()) == tot and tot > 0
    ph, mk = eng.mf_decode(_t(stack), black_thr=40)
    xyz2, valid2, k2, n2 = eng.match_triangulate_phase(ph, mk)
    assert (k2.cpu().numpy() == k.cpu().numpy()).all() and (bits(xyz2.cpu().numpy()) == bits(xyz.cpu().numpy())).all()
    h_xyz = np.empty((2, H, W, 3), np.float32)
    h_valid = np.empty((2, H, W), np.uint8)
    h_k = np.empty((2, H, W), np.int32)
    n_host = eng.run_mf_host(stack, h_xyz, h_valid, h_k)
    assert n_host == tot and (h_k == k.cpu().numpy()).all() and (bits(h_xyz) == bits(xyz.cpu().numpy())).all()
    # ---- GE: fused entry with colour, and the code match entry point ----
    nc = oracle.gray_num_bits(W)
    g = synth.synth_gray(W, H, seed=8, integer_disparity=False, noise_dn=2.0)[None]
    xyz, valid, k, colr, n = eng.run_ge(_t(g), nc, black_thr=40, white_thr=3, have_color=True)
    cols, mks = zip(*[(lambda r: (r[0], r[2]))(oracle.gray_decode(g[0, cam], nc, 0, 40, 3, W, H)) for cam in range(2)])
    xyz_o, valid_o, k_o, col_o, n_o = oracle.ge_triangulate(cols[0], mks[0], cols[1], mks[1], Q, whiteL=g[0, 0, 0], whiteR=g[0, 1, 0])
    assert (k[0].cpu().numpy() == k_o).all() and (valid[0].cpu().numpy() == valid_o).all()
    assert (colr[0].cpu().numpy() == col_o).all() and (bits(xyz[0].cpu().numpy()) == bits(xyz_o)).all()
    assert int(n.item()) == n_o and n_o > 0
    col, _, gm = eng.gray_decode(_t(g), nc, 0, 40, 3, W, H)
    xyz3, valid3, k3, _, n3 = eng.match_triangulate_code(col, gm)
    assert (k3[0].cpu().numpy() == k_o).all() and (bits(xyz3[0].cpu().numpy()) == bits(xyz_o)).all()


# ------------------------------------------------------------------------------------------------
# fuzz: pure-noise and low-entropy stacks (stress the hash table, the chains, duplicate handling, thresholds)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", list(range(8)))
def test_fuzz_random_stacks_match_the_oracle(cuda_engine_factory, oracle, seed):
    """Random image bytes are the opposite of a scene: phases / codes are uncorrelated between cameras, values collide
    and repeat at random.  Every pipeline must still reproduce the oracle bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    W = int(rng.choice([16, 48, 64, 160, 256, 320, 1280]))
    H = int(rng.integers(1, 6))
    levels = int(rng.choice([2, 3, 8, 256]))          # few grey levels -> many equal phases, degenerate (0/0) pixels
    black_thr = int(rng.choice([0, 5, 40, 120]))
    eng = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    rigid = None
    if seed % 2:
        rigid = np.array([[0.98, -0.17, 0.05, 12.5], [0.17, 0.98, 0.02, -3.25], [-0.05, -0.01, 0.99, 40.0]], np.float32)
    eng.set_calib(cams, Q, rigid)

    def noise(n):
        st = (rng.integers(0, levels, (1, 2, n, H, W)) * (255 // max(levels - 1, 1))).astype(np.uint8)
        st[:, :, 0] = rng.integers(100, 256, (1, 2, H, W))       # white / black so that part of the image is lit
        st[:, :, 1] = rng.integers(0, 120, (1, 2, H, W))
        return st

    mf = noise(14)
    xyz, valid, k, n = eng.run_mf(_t(mf), black_thr=black_thr)
    xyz_o, valid_o, k_o, n_o = oracle.run_mf(mf[0], cams, Q, black_thr=black_thr, rigid=rigid)
    assert (k[0].cpu().numpy() == k_o).all() and (valid[0].cpu().numpy() == valid_o).all()
    assert (bits(xyz[0].cpu().numpy())[valid_o > 0] == bits(xyz_o)[valid_o > 0]).all() and int(n.item()) == n_o

    nc = oracle.gray_num_bits(W)
    white_thr = int(rng.choice([0, 1, 30, 200]))
    ge = noise(2 + 2 * nc)
    xyz, valid, k, colr, n = eng.run_ge(_t(ge), nc, black_thr=black_thr, white_thr=white_thr, have_color=True)
    cols, mks = zip(*[(lambda r: (r[0], r[2]))(oracle.gray_decode(ge[0, cam], nc, 0, black_thr, white_thr, W, H)) for cam in range(2)])
    xyz_o, valid_o, k_o, col_o, n_o = oracle.ge_triangulate(cols[0], mks[0], cols[1], mks[1], Q, rigid=rigid, whiteL=ge[0, 0, 0], whiteR=ge[0, 1, 0])
    assert (k[0].cpu().numpy() == k_o).all() and (valid[0].cpu().numpy() == valid_o).all()
    assert (colr[0].cpu().numpy() == col_o).all()
    assert (bits(xyz[0].cpu().numpy())[valid_o > 0] == bits(xyz_o)[valid_o > 0]).all() and int(n.item()) == n_o
