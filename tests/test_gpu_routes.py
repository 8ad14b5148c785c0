"""GPU parity tests (-m gpu) for the launch routes that round 1 only compared with themselves: every test here
compares the CUDA path (through the C ABI) with the CPU ORACLE, never with another CUDA kernel.

  * BASELINE config 4's route: 2048-wide rows (one wide CTA per SM in the fused kernel), strict mode;
  * BASELINE config 5's route: 4096-wide rows, 4 frequencies x 8 shifts, corrected mode (K1 4x8 + wide-row match);
  * a full 1280x1024 frame with sub-pixel disparity and sensor noise (the dedupe tables see ~2x the distinct
    phase values of the noise-free case);
  * Gray-only bucket triangulation at 640x480 (BASELINE config 1's resolution);
  * device pointers that are not 16-byte aligned (staged through aligned copies by the library).
"""
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


def _assert_cloud_equal(xyz, valid, k, n, xyz_o, valid_o, k_o, n_o, what):
    k, valid, xyz = k.cpu().numpy(), valid.cpu().numpy(), xyz.cpu().numpy()
    assert (k == k_o).all(), f"{what}: {(k != k_o).sum()} match columns differ"
    assert (valid == valid_o).all(), f"{what}: valid flags differ"
    assert (np.isnan(xyz) == np.isnan(xyz_o)).all(), f"{what}: NaN pattern differs"
    ok = ~np.isnan(xyz_o)
    assert np.allclose(xyz[ok], xyz_o[ok], rtol=RTOL, atol=1e-6), f"{what}: XYZ outside 1e-4 relative"
    assert (bits(xyz) == bits(xyz_o)).all(), f"{what}: XYZ not bit-exact"
    assert n == n_o, f"{what}: point count {n} != {n_o}"


# ------------------------------------------------------------------------------------------------
# config 4's route: W = 2048 (Duke/mfreconstruct.cpp:160-334 at a width the reference's GUI cannot select)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("intd,noise,rigid", [(True, 0.0, False), (False, 2.0, True)])
def test_run_mf_2048_wide_strict_vs_oracle(cuda_engine_factory, oracle, intd, noise, rigid):
    W, H, B = 2048, 48, 2
    eng = cuda_engine_factory(W, H, B)
    cams, Q = slr_b200.synthetic_rig(W, H)
    M = np.array([[0.98, -0.17, 0.05, 12.5], [0.17, 0.98, 0.02, -3.25], [-0.05, -0.01, 0.99, 40.0]], np.float32) \
        if rigid else None
    eng.set_calib(cams, Q, M)
    stack = np.stack([synth.synth_mf(W, H, seed=s, integer_disparity=intd, noise_dn=noise) for s in (21, 22)])
    xyz, valid, k, n = eng.run_mf(_t(stack), black_thr=40)
    tot = 0
    for b in range(B):
        xyz_o, valid_o, k_o, n_o = oracle.run_mf(stack[b], cams, Q, rigid=M, nthreads=oracle.max_threads())
        _assert_cloud_equal(xyz[b], valid[b], k[b], n_o, xyz_o, valid_o, k_o, n_o, f"2048-wide scan {b}")
        tot += n_o
    assert int(n.item()) == tot and tot > 0.3 * B * W * H


def test_match_phase_2048_wide_vs_oracle(cuda_engine_factory, oracle):
    """slr_match_triangulate_phase on 2048-wide rows: oracle-decoded phases in, oracle match out."""
    W, H = 2048, 24
    eng = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = synth.synth_mf(W, H, seed=31, integer_disparity=False, noise_dn=1.0)
    ph = np.empty((1, 2, H, W), np.float32)
    mk = np.empty((1, 2, H, W), np.uint8)
    for cam in range(2):
        ph[0, cam], mk[0, cam] = oracle.mf_decode(stack[cam], black_thr=40)
    xyz, valid, k, n = eng.match_triangulate_phase(_t(ph), _t(mk))
    xyz_o, valid_o, k_o, n_o = oracle.mf_triangulate(ph[0, 0], mk[0, 0], ph[0, 1], mk[0, 1], cams, Q,
                                                     nthreads=oracle.max_threads())
    _assert_cloud_equal(xyz[0], valid[0], k[0], int(n.item()), xyz_o, valid_o, k_o, n_o, "2048-wide phase match")
    assert n_o > 0.3 * W * H


# ------------------------------------------------------------------------------------------------
# config 5's route: 4096 wide, F = 4, S = 8, corrected mode.  Corrected mode has no reference counterpart
# (SURVEY.md §0 F8): every comparison below is ORACLE-ONLY.
# ------------------------------------------------------------------------------------------------
def test_run_mf_config5_route_vs_oracle(cuda_engine_factory, oracle):
    W, H, F, S = 4096, 16, 4, 8
    eng = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = synth.synth_mf(W, H, seed=51, integer_disparity=False, noise_dn=1.0, F=F, S=S)[None]
    st = _t(stack)
    # (1) decode: within the north star's 1e-4 (phase is circular with period 255)
    ph, mk = eng.mf_decode(st, F=F, S=S, black_thr=40, mode=slr_b200.MODE_CORRECTED)
    ph, mk = ph.cpu().numpy(), mk.cpu().numpy()
    for cam in range(2):
        ph_o, mk_o = oracle.mf_decode(stack[0, cam], F=F, S=S, black_thr=40, mode=1, nthreads=oracle.max_threads())
        assert (mk[0, cam] == mk_o).all()
        ok = mk_o == 1
        d = np.abs(ph[0, cam][ok] - ph_o[ok])
        d = np.minimum(d, 255.0 - d)
        assert (d <= RTOL * 255.0).all(), d.max()
    # (2) match + triangulate of the whole pipeline == the oracle's match on the very phases the GPU decoded: exact
    xyz, valid, k, n = eng.run_mf(st, F=F, S=S, black_thr=40, mode=slr_b200.MODE_CORRECTED)
    xyz_o, valid_o, k_o, n_o = oracle.mf_triangulate(ph[0, 0], mk[0, 0], ph[0, 1], mk[0, 1], cams, Q,
                                                     nthreads=oracle.max_threads())
    _assert_cloud_equal(xyz[0], valid[0], k[0], int(n.item()), xyz_o, valid_o, k_o, n_o, "config-5 route, GPU phases")
    assert n_o > 0.3 * W * H
    # (3) end to end against oracle.run_mf(mode=1): a phase difference of 1e-5 can move a |pL - pR| < 0.1 decision,
    # so match columns agree for all but a sliver of borderline pixels; where they agree XYZ is within 1e-4
    xyz_e, valid_e, k_e, n_e = oracle.run_mf(stack[0], cams, Q, F=F, S=S, black_thr=40, mode=1,
                                             nthreads=oracle.max_threads())
    kg = k[0].cpu().numpy()
    same = kg == k_e
    assert same.mean() > 0.995, same.mean()
    sel = same & (k_e >= 0)
    g = xyz[0].cpu().numpy()[sel]
    assert np.allclose(g, xyz_e[sel], rtol=RTOL, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# full frame, sub-pixel disparity + noise (sigma in {0, 2} DN)
# ------------------------------------------------------------------------------------------------
def test_run_mf_full_frame_noisy_vs_oracle(cuda_engine_factory, oracle):
    W, H = 1280, 1024
    eng = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = eng.synth_mf(1, seed=43, integer_disparity=False, noise_dn=2.0)
    xyz, valid, k, n = eng.run_mf(stack, black_thr=40)
    hs = stack.cpu().numpy()
    xyz_o, valid_o, k_o, n_o = oracle.run_mf(hs[0], cams, Q, nthreads=oracle.max_threads())
    _assert_cloud_equal(xyz[0], valid[0], k[0], int(n.item()), xyz_o, valid_o, k_o, n_o, "noisy full frame")
    assert n_o > 0.2 * W * H
    # size-independent property: a second run reproduces the first bit for bit
    xyz2, valid2, k2, n2 = eng.run_mf(stack, black_thr=40)
    assert (bits(xyz2.cpu().numpy()) == bits(xyz.cpu().numpy())).all() and int(n2.item()) == int(valid2.sum().item())


# ------------------------------------------------------------------------------------------------
# Gray-only buckets at 640x480 (Duke/reconstruct.cpp:56-74, 417-481)
# ------------------------------------------------------------------------------------------------
def test_k3c_bucket_triangulate_640x480_vs_oracle(cuda_engine_factory, oracle):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import cases
    W, H = 640, 480
    eng = cuda_engine_factory(W, H, 1)
    cams = cases.gray_only_rig(W, H)
    _, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q, cases.RIGID)
    stack = synth.synth_gray(W, H, seed=61, noise_dn=2.0, rows=True, integer_disparity=True)[None]
    nc, nr = oracle.gray_num_bits(W), oracle.gray_num_bits(H)
    col, row, mk = eng.gray_decode(_t(stack), nc, nr, black_thr=40, white_thr=3, scan_w=W, scan_h=H)
    ssum, cnt, n = eng.bucket_triangulate(col, row, mk, W, H)
    d = [oracle.gray_decode(stack[0, cam], nc, nr, 40, 3, W, H) for cam in range(2)]
    for cam in range(2):
        assert (col[0, cam].cpu().numpy() == d[cam][0]).all() and (row[0, cam].cpu().numpy() == d[cam][1]).all()
        assert (mk[0, cam].cpu().numpy() == d[cam][2]).all()
    s_o, c_o, n_o = oracle.gray_triangulate(d[0][0], d[0][1], d[0][2], d[1][0], d[1][1], d[1][2], W, H, cams, cases.RIGID)
    assert (cnt[0].cpu().numpy() == c_o).all()
    g = ssum[0].cpu().numpy()
    assert (bits(g[c_o > 0]) == bits(s_o[c_o > 0])).all()
    assert int(n.item()) == n_o and n_o > 1000


# ------------------------------------------------------------------------------------------------
# device pointers off the 16-byte grid (the fast paths need aligned bases; others are staged)
# ------------------------------------------------------------------------------------------------
def _offset_like(t, off_bytes):
    """A tensor with t's contents whose data pointer is t-aligned + off_bytes (off_bytes a multiple of the item size)."""
    import torch
    raw = torch.empty(t.numel() * t.element_size() + 64, dtype=torch.uint8, device=t.device)
    base = (-raw.data_ptr()) % 16 + off_bytes
    view = raw[base:base + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
    view.copy_(t)
    assert view.data_ptr() % 16 == off_bytes % 16
    return view


def test_unaligned_device_pointers_match_the_oracle(cuda_engine_factory, oracle):
    import torch
    W, H = 320, 12
    eng = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = synth.synth_mf(W, H, seed=71, noise_dn=1.0)[None]
    xyz_o, valid_o, k_o, n_o = oracle.run_mf(stack[0], cams, Q)
    # MF pipeline: input stack 4 bytes off, outputs 4 / 1 / 4 bytes off
    st = _offset_like(_t(stack), 4)
    out = (_offset_like(torch.zeros((1, H, W, 3), dtype=torch.float32, device="cuda"), 4),
           _offset_like(torch.zeros((1, H, W), dtype=torch.uint8, device="cuda"), 1),
           _offset_like(torch.zeros((1, H, W), dtype=torch.int32, device="cuda"), 4), None,
           torch.zeros(1, dtype=torch.int64, device="cuda"))
    xyz, valid, k, n = eng.run_mf(st, black_thr=40, out=out)
    _assert_cloud_equal(xyz[0], valid[0], k[0], int(n.item()), xyz_o, valid_o, k_o, n_o, "unaligned run_mf")
    # phase-match entry point with unaligned phase rows
    ph = np.empty((1, 2, H, W), np.float32)
    mk = np.empty((1, 2, H, W), np.uint8)
    for cam in range(2):
        ph[0, cam], mk[0, cam] = oracle.mf_decode(stack[0, cam], black_thr=40)
    xyz, valid, k, n = eng.match_triangulate_phase(_offset_like(_t(ph), 8), _offset_like(_t(mk), 3))
    _assert_cloud_equal(xyz[0], valid[0], k[0], int(n.item()), xyz_o, valid_o, k_o, n_o, "unaligned phase match")
    # Gray-EPI pipeline
    nc = oracle.gray_num_bits(W)
    g = synth.synth_gray(W, H, seed=72, integer_disparity=False, noise_dn=2.0)[None]
    xyz, valid, k, colr, n = eng.run_ge(_offset_like(_t(g), 2), nc, black_thr=40, white_thr=3, have_color=False)
    cols, mks = zip(*[(lambda r: (r[0], r[2]))(oracle.gray_decode(g[0, cam], nc, 0, 40, 3, W, H)) for cam in range(2)])
    xyz_g, valid_g, k_g, _, n_g = oracle.ge_triangulate(cols[0], mks[0], cols[1], mks[1], Q)
    _assert_cloud_equal(xyz[0], valid[0], k[0], int(n.item()), xyz_g, valid_g, k_g, n_g, "unaligned run_ge")
    col = np.stack(cols)[None]
    gm = np.stack(mks)[None]
    xyz, valid, k, _, n = eng.match_triangulate_code(_offset_like(_t(col), 4), _offset_like(_t(gm), 5))
    _assert_cloud_equal(xyz[0], valid[0], k[0], int(n.item()), xyz_g, valid_g, k_g, n_g, "unaligned code match")


# ------------------------------------------------------------------------------------------------
# row bands: an engine of the band's height with slr_set_row_offset == the same rows of the full image
# ------------------------------------------------------------------------------------------------
def test_row_band_engines_reproduce_the_full_image(cuda_engine_factory):
    import torch
    W, H = 640, 60
    full = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    full.set_calib(cams, Q)
    mf = full.synth_mf(1, seed=81, integer_disparity=False, noise_dn=1.0)
    ge = full.synth_gray(1, seed=82, integer_disparity=False, noise_dn=2.0)
    nb = slr_b200.gray_num_bits(W)
    xyz_f, valid_f, k_f, _ = full.run_mf(mf, black_thr=40)
    gx_f, gv_f, gk_f, _, _ = full.run_ge(ge, nb, black_thr=40, white_thr=3)
    for lo, hi in ((0, 20), (20, 47), (47, 60)):
        band = cuda_engine_factory(W, hi - lo, 1)
        band.set_calib(cams, Q)
        band.set_row_offset(lo)
        xyz, valid, k, _ = band.run_mf(mf[:, :, :, lo:hi].contiguous(), black_thr=40)
        assert torch.equal(k, k_f[:, lo:hi]) and torch.equal(valid, valid_f[:, lo:hi])
        assert (bits(xyz.cpu().numpy()) == bits(xyz_f[:, lo:hi].cpu().numpy())).all()
        gx, gv, gk, _, _ = band.run_ge(ge[:, :, :, lo:hi].contiguous(), nb, black_thr=40, white_thr=3)
        assert torch.equal(gk, gk_f[:, lo:hi]) and torch.equal(gv, gv_f[:, lo:hi])
        assert (bits(gx.cpu().numpy()) == bits(gx_f[:, lo:hi].cpu().numpy())).all()


# ------------------------------------------------------------------------------------------------
# SURVEY.md 8f N1 as written: rectification inside the fused kernel's stage fill (slr_run_mf_raw) —
# stereoRect::doStereoRectify (Duke/stereorect.cpp:26-34) + MFReconstruct::runReconstruction, raw bytes -> XYZ
# ------------------------------------------------------------------------------------------------
def _rect_maps(W, H, seed, border=False):
    cv2 = pytest.importorskip("cv2")
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    maps1, maps2 = [], []
    for cam in range(2):
        mx = xs * (1.004 - 0.008 * cam) - 2 + 1.5 * np.sin(ys / 17 + seed)
        my = ys * 0.995 + 1 - 2 * cam + 1.5 * np.cos(xs / 41 + seed)       # rows drift: groups straddle source rows
        if border:                                                         # taps outside the image: BORDER_CONSTANT 0
            mx, my = mx - 6, my - 3
        mx[:2], my[:2] = xs[:2], ys[:2]                                    # integer coordinates: the (0,0) weight entry
        m1, m2 = cv2.convertMaps(mx, my, cv2.CV_16SC2)
        maps1.append(m1)
        maps2.append(m2)
    return np.stack(maps1), np.stack(maps2)


@pytest.mark.parametrize("W,H,B,mode,border", [(1280, 40, 3, slr_b200.MODE_STRICT, False),
                                               (640, 33, 2, slr_b200.MODE_STRICT, True),
                                               (1280, 24, 2, slr_b200.MODE_CORRECTED, False),
                                               (2048, 20, 1, slr_b200.MODE_STRICT, False)])
def test_run_mf_raw_rectifies_inside_the_fused_kernel_vs_oracle(cuda_engine_factory, oracle, W, H, B, mode, border):
    eng = cuda_engine_factory(W, H, B)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    m1, m2 = _rect_maps(W, H, 0.7, border)
    eng.set_rectify_maps(m1, m2)
    raw = np.stack([synth.synth_mf(W, H, seed=40 + s, integer_disparity=False, noise_dn=1.5) for s in range(B)])
    l0 = eng.launches()
    xyz, valid, k, n = eng.run_mf_raw(_t(raw), black_thr=40, mode=mode)
    import torch
    torch.cuda.synchronize()
    if W <= 1280:    # (2048-wide rows exceed the dataflow kernel's shared memory: K0 + the wide-row kernel per scan)
        assert eng.launches() - l0 == 1, "raw stacks -> XYZ must be ONE kernel"
    tot = 0
    for b in range(B):
        rect = np.stack([[oracle.remap_linear(raw[b, cam, i], m1[cam], m2[cam]) for i in range(raw.shape[2])]
                         for cam in range(2)])
        xyz_o, valid_o, k_o, n_o = oracle.run_mf(rect, cams, Q, mode=mode, nthreads=oracle.max_threads())
        if mode == slr_b200.MODE_STRICT:
            _assert_cloud_equal(xyz[b], valid[b], k[b], n_o, xyz_o, valid_o, k_o, n_o, f"raw scan {b}")
        else:   # corrected mode (oracle-only): phases agree to 1e-4, matches may differ where a phase sits on the window edge
            same = k[b].cpu().numpy() == k_o
            assert same.mean() > 0.995, same.mean()
            ok = same & (valid_o != 0)
            assert np.allclose(xyz[b].cpu().numpy()[ok], xyz_o[ok], rtol=RTOL, atol=1e-5)
        tot += n_o
    assert tot > 1000
    if mode == slr_b200.MODE_STRICT:
        assert int(n.item()) == tot
    # the two-kernel route (K0 into HBM, then the fused kernel) gives the same bits
    xyz2, valid2, k2, n2 = eng.run_mf(eng.rectify_stack(_t(raw)), black_thr=40, mode=mode)
    assert (bits(xyz2.cpu().numpy()) == bits(xyz.cpu().numpy())).all() and (k2 == k).all() and int(n2.item()) == int(n.item())


# ------------------------------------------------------------------------------------------------
# PNG ingest — scanlines as the zlib stream holds them (filter byte + filtered row) are unfiltered on
# the GPU, image by image; the cloud comes back in the reference's PointCloudImage storage layout
# ------------------------------------------------------------------------------------------------
def _png_filter_rows(img, types):
    """forward PNG filters None / Sub / Up of an 8-bit grey image -> [H, 1 + W] scanlines"""
    H, W = img.shape
    out = np.empty((H, W + 1), np.uint8)
    a = img.astype(np.int16)
    for y in range(H):
        t = int(types[y])
        if t == 0:
            f = a[y]
        elif t == 1:
            f = a[y] - np.concatenate([[0], a[y, :-1]])
        else:
            f = a[y] - (a[y - 1] if y else 0)
        out[y, 0] = t
        out[y, 1:] = (f & 255).astype(np.uint8)
    return out


@pytest.mark.parametrize("W,H,scan_w,scan_h,raw", [(1280, 48, 1280, 1024, False), (640, 40, 30, 500, True)])
def test_png_ingest_unfilters_on_the_gpu_and_returns_the_pointcloudimage_layout(cuda_engine_factory, oracle, W, H,
                                                                                 scan_w, scan_h, raw):
    eng = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    if raw:
        m1, m2 = _rect_maps(W, H, 0.2)
        eng.set_rectify_maps(m1, m2)
        eng.set_host_input_raw(True)
    stack = synth.synth_mf(W, H, seed=91, integer_disparity=False, noise_dn=2.0)      # [2, 14, H, W]
    rng = np.random.default_rng(5)
    images, filtered = [], []
    for k in range(28):
        img = stack[k // 14, k % 14]
        kind = k % 4
        if kind == 0:
            images.append(img), filtered.append(False)                                  # decoded by the caller
        else:
            types = np.ones(H, np.int64) if kind == 1 else rng.integers(0, 3 if kind == 2 else 2, H)
            images.append(_png_filter_rows(img, types)), filtered.append(True)
    h_xyz = np.empty((H, W, 3), np.float32)
    h_valid = np.empty((H, W), np.uint8)
    h_sum = np.full((scan_h, scan_w, 3), -7.0, np.float32)
    h_cnt = np.full((scan_h, scan_w), 9, np.uint8)
    n = eng.run_mf_ingested(images, filtered, h_sum, h_cnt, h_xyz, h_valid, scan_w, scan_h)
    # the same scan through the device-pointer entry points
    if raw:
        xyz, valid, k, n2 = eng.run_mf_raw(_t(stack[None]))
        eng.set_host_input_raw(False)
        rect = np.stack([[oracle.remap_linear(stack[cam, i], m1[cam], m2[cam]) for i in range(14)] for cam in range(2)])
    else:
        xyz, valid, k, n2 = eng.run_mf(_t(stack[None]))
        rect = stack
    assert n == int(n2.item()) and n > 1000
    assert (bits(h_xyz) == bits(xyz[0].cpu().numpy())).all() and (h_valid == valid[0].cpu().numpy()).all()
    xyz_o, valid_o, k_o, n_o = oracle.run_mf(rect, cams, Q, nthreads=oracle.max_threads())
    assert n == n_o and (bits(h_xyz) == bits(xyz_o)).all()
    # PointCloudImage(scan_w, scan_h) after addPoint(i, j, p) for every valid (row i, column j): cell (i_w = i, j_h = j),
    # stored at [j_h][i_w]; i >= scan_w or j >= scan_h dropped (Duke/pointcloudimage.cpp:86-97)
    exp_sum = np.zeros((scan_h, scan_w, 3), np.float32)
    exp_cnt = np.zeros((scan_h, scan_w), np.uint8)
    hh, ww = min(H, scan_w), min(W, scan_h)
    v = valid_o[:hh, :ww] != 0
    exp_cnt[:ww, :hh] = v.T
    exp_sum[:ww, :hh] = np.where(v[..., None], xyz_o[:hh, :ww], 0).transpose(1, 0, 2)
    assert (h_cnt == exp_cnt).all() and (bits(h_sum) == bits(exp_sum)).all()


# ------------------------------------------------------------------------------------------------
# multi-scan merge — the clouds of several scans, each with its transfer matrix, as one point list
# ------------------------------------------------------------------------------------------------
def test_merge_scans_equals_the_per_scan_rigid_epilogue_and_the_oracle(cuda_engine_factory, oracle):
    W, H, B = 640, 40, 3
    eng = cuda_engine_factory(W, H, B)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = np.stack([synth.synth_mf(W, H, seed=70 + s, integer_disparity=False, noise_dn=1.0) for s in range(B)])
    Ms = np.array([np.eye(3, 4),
                   [[0.98, -0.17, 0.05, 12.5], [0.17, 0.98, 0.02, -3.25], [-0.05, -0.01, 0.99, 40.0]],
                   [[0.7, 0.1, -0.7, -100.0], [0.0, 0.99, 0.14, 3.0], [0.71, -0.1, 0.69, 250.5]]], np.float32)
    has = np.array([0, 1, 1], np.uint8)                 # scan 0 is the common frame (scanSN == 0: no transfer matrix)
    xyz, valid, k, n = eng.run_mf(_t(stack))
    pts, src = eng.merge_scans(xyz, valid, Ms, has)
    pts, src = pts.cpu().numpy(), src.cpu().numpy()
    exp_pts, exp_src = [], []
    for s in range(B):
        xyz_o, valid_o, _, n_o = oracle.run_mf(stack[s], cams, Q, rigid=Ms[s] if has[s] else None,
                                               nthreads=oracle.max_threads())
        idx = np.flatnonzero(valid_o.reshape(-1))
        exp_pts.append(xyz_o.reshape(-1, 3)[idx])
        exp_src.append(idx + s * W * H)
    exp_pts, exp_src = np.concatenate(exp_pts), np.concatenate(exp_src)
    assert len(pts) == len(exp_pts) == int(n.item()) and len(pts) > 3000
    assert (src == exp_src).all() and (bits(pts) == bits(exp_pts)).all()
    # no matrices: plain ordered compaction
    p0, s0 = eng.merge_scans(xyz, valid)
    assert (s0.cpu().numpy() == exp_src).all()
    assert (bits(p0.cpu().numpy()) == bits(xyz.cpu().numpy().reshape(-1, 3)[exp_src])).all()


def test_raw_and_merge_edge_cases(cuda_engine_factory, oracle):
    """slr_run_mf_raw on a width the TMA rows cannot take (100: K0's per-pixel kernel, then the padded route), maps that
    send whole rows outside the image, a merge without a single valid point, and argument errors."""
    import torch
    W, H = 100, 20
    eng = cuda_engine_factory(W, H, 2)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    with pytest.raises(slr_b200.SlrError):                 # no maps yet
        eng.run_mf_raw(_t(np.zeros((1, 2, 14, H, W), np.uint8)))
    m1, m2 = _rect_maps(W, H, 0.4)
    m1[1, 5:8, :, 1] = -30                                 # right camera, rows 5..7: every tap above the image
    eng.set_rectify_maps(m1, m2)
    raw = np.stack([synth.synth_mf(W, H, seed=80 + s, noise_dn=1.0) for s in range(2)])
    xyz, valid, k, n = eng.run_mf_raw(_t(raw))
    tot = 0
    for b in range(2):
        rect = np.stack([[oracle.remap_linear(raw[b, cam, i], m1[cam], m2[cam]) for i in range(14)] for cam in range(2)])
        assert (rect[1, :, 5:8] == 0).all()
        xyz_o, valid_o, k_o, n_o = oracle.run_mf(rect, cams, Q)
        _assert_cloud_equal(xyz[b], valid[b], k[b], n_o, xyz_o, valid_o, k_o, n_o, f"ragged raw scan {b}")
        assert (valid_o[5:8] == 0).all()
        tot += n_o
    assert int(n.item()) == tot and tot > 100
    # merge: nothing valid -> an empty list; everything valid -> identity order
    none = torch.zeros((2, H, W), dtype=torch.uint8, device="cuda")
    pts, src = eng.merge_scans(xyz, none)
    assert pts.shape[0] == 0 and src.shape[0] == 0
    allv = torch.ones((2, H, W), dtype=torch.uint8, device="cuda")
    dense = torch.arange(2 * H * W * 3, dtype=torch.float32, device="cuda").reshape(2, H, W, 3)
    pts, src = eng.merge_scans(dense, allv)
    assert (src.cpu().numpy() == np.arange(2 * H * W)).all() and torch.equal(pts, dense.reshape(-1, 3))
    # ingest: odd image counts and a stack of the wrong size are refused
    with pytest.raises(slr_b200.SlrError):
        eng.run_mf_ingested([np.zeros((H, W), np.uint8)] * 3, [False] * 3, h_xyz=np.empty((H, W, 3), np.float32),
                            h_valid=np.empty((H, W), np.uint8))
    with pytest.raises(slr_b200.SlrError):
        eng.run_mf_ingested([np.zeros((H, W), np.uint8)] * 4, [False] * 4, h_xyz=np.empty((H, W, 3), np.float32),
                            h_valid=np.empty((H, W), np.uint8))


@pytest.mark.parametrize("black_thr", [-1, 0, 179, 180, 254, 300])
def test_run_mf_extreme_shadow_thresholds_vs_oracle(cuda_engine_factory, oracle, black_thr):
    """computeShadows' threshold at its extremes (Duke/mfreconstruct.cpp:199-204): below zero every pixel is decoded
    (degenerate G1 == G3 && G4 == G2 pixels included: the oracle's 'pixel dropped'), at white - black (180 in the synthetic
    scene) and above nothing is lit and the cloud is empty."""
    W, H = 640, 24
    eng = cuda_engine_factory(W, H, 1)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = synth.synth_mf(W, H, seed=33, integer_disparity=False, noise_dn=0.0)
    stack[:, 2:, :, 100:140] = 77                      # a flat patch: every frequency degenerate
    xyz, valid, k, n = eng.run_mf(_t(stack[None]), black_thr=black_thr)
    xyz_o, valid_o, k_o, n_o = oracle.run_mf(stack, cams, Q, black_thr=black_thr)
    _assert_cloud_equal(xyz[0], valid[0], k[0], int(n.item()), xyz_o, valid_o, k_o, n_o, f"black_thr {black_thr}")
    if black_thr >= 180:
        assert n_o == 0
    if black_thr < 0:
        assert n_o > 0 and (valid_o[:, 100:140] == 0).all()


@pytest.mark.parametrize("W,H,mode", [(1536, 40, slr_b200.MODE_STRICT), (1920, 33, slr_b200.MODE_STRICT),
                                      (2048, 24, slr_b200.MODE_CORRECTED), (1424, 17, slr_b200.MODE_STRICT)])
def test_widths_between_the_dataflow_and_the_wide_row_kernel_vs_oracle(cuda_engine_factory, oracle, W, H, mode):
    """Rows too wide for the dataflow kernel's four row contexts (more than 1408 pixels) run k_fused_mf's one 1024-thread
    CTA per SM: widths around the hand-over, partially filled tasks, both decode modes.  (A two-context dataflow schedule
    for these widths was built and measured 10 % slower than k_fused_mf at 2048 wide; not kept.)"""
    B = 3
    eng = cuda_engine_factory(W, H, B)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = np.stack([synth.synth_mf(W, H, seed=50 + s, integer_disparity=False, noise_dn=1.5) for s in range(B)])
    l0 = eng.launches()
    xyz, valid, k, n = eng.run_mf(_t(stack), black_thr=40, mode=mode)
    assert eng.launches() - l0 == 1
    tot = 0
    for b in range(B):
        xyz_o, valid_o, k_o, n_o = oracle.run_mf(stack[b], cams, Q, mode=mode, nthreads=oracle.max_threads())
        if mode == slr_b200.MODE_STRICT:
            _assert_cloud_equal(xyz[b], valid[b], k[b], n_o, xyz_o, valid_o, k_o, n_o, f"{W}-wide scan {b}")
        else:
            same = k[b].cpu().numpy() == k_o
            assert same.mean() > 0.995, same.mean()
            ok = same & (valid_o != 0)
            assert np.allclose(xyz[b].cpu().numpy()[ok], xyz_o[ok], rtol=RTOL, atol=1e-5)
        tot += n_o
    assert tot > 1000
    if mode == slr_b200.MODE_STRICT:
        assert int(n.item()) == tot
