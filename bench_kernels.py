#!/usr/bin/env python
"""bench_kernels.py — per-kernel measurements for every row of SURVEY.md §8 (not the driver's contract; that is
bench.py).  For each entry point: CUDA-event time on device-resident inputs, algorithmic bytes,
achieved GB/s against the measured HBM peak, and the oracle port of the same reference loop on the host cores.

    python bench_kernels.py [--batch 8] [--reps 10] [--no-cpu]   ->  markdown table on stdout, JSON lines on stderr
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import numpy as np
    import torch
    import slr_b200
    from slr_b200 import synth
    from bench import load_peak

    assert torch.cuda.is_available(), "needs a CUDA device (no CPU fallback)"
    peak, peak_src = load_peak()
    B = args.batch
    nthreads = len(os.sched_getaffinity(0))
    orc = None
    if not args.no_cpu:
        import oracle_lib
        orc = oracle_lib.load()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for a, b in evs:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        return ts[len(ts) // 2]

    def cpu_time(fn, min_s=2.0):
        if orc is None:
            return None
        t0 = time.perf_counter()
        n = 0
        while True:
            fn()
            n += 1
            dt = time.perf_counter() - t0
            if dt > min_s or n >= 20:
                return dt / n

    rows = []

    def report(name, replaces, ms, alg_bytes, cpu_s, units, unit_name):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        row = {"kernel": name, "replaces": replaces, "ms": ms, "algorithmic_MB": alg_bytes / 1e6, "GB/s": gbs, "frac_of_hbm": gbs / peak,
               unit_name + "/s": units / (ms * 1e-3), "cpu_port_s_per_scan": cpu_s, "cpu_threads": nthreads,
               "speedup_vs_cpu_port": (cpu_s * B / (ms * 1e-3)) if cpu_s else None}
        rows.append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)

    # ---------------- 1280x1024 multi-frequency ----------------
    W, H = 1280, 1024
    P = W * H
    eng = slr_b200.Engine(W, H, max_batch=B)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    h_mf = np.stack([synth.synth_mf(W, H, seed=1 + s) for s in range(min(B, 4))])
    mf = torch.from_numpy(h_mf).cuda().repeat((B + 3) // 4, 1, 1, 1, 1)[:B].contiguous()
    ph, mk = eng.mf_decode(mf)
    out = eng._outputs(B, want_k=False)

    ms = timed(lambda: eng.mf_decode(mf))
    c = cpu_time(lambda: [orc.mf_decode(h_mf[0, cam], nthreads=nthreads) for cam in range(2)]) if orc else None
    report("k1_mf_decode (strict)", "computeShadows+decodePatterns+getPhase", ms, B * 2 * P * (14 + 5), c, B * 2 * P, "pixels")

    ms = timed(lambda: eng.mf_decode(mf, mode=slr_b200.MODE_CORRECTED))
    c = cpu_time(lambda: [orc.mf_decode(h_mf[0, cam], mode=1, nthreads=nthreads) for cam in range(2)]) if orc else None
    report("k1_mf_decode (corrected: atan2 + heterodyne, BASELINE config 2)", "(no reference counterpart)", ms, B * 2 * P * (14 + 5), c, B * 2 * P, "pixels")

    ms = timed(lambda: eng.match_triangulate_phase(ph, mk, want_k=False))
    ph_h2, mk_h2 = ph[0].cpu().numpy(), mk[0].cpu().numpy()
    c = cpu_time(lambda: orc.mf_triangulate(ph_h2[0], mk_h2[0], ph_h2[1], mk_h2[1], cams, Q, nthreads=nthreads)) if orc else None
    report("k3a phase match+triangulate", "MFReconstruct::triangulation", ms, B * (2 * P * 5 + P * 13), c, B * P, "pixels")

    ms = timed(lambda: eng.run_mf(mf, out=out))
    c = cpu_time(lambda: orc.run_mf(h_mf[0], cams, Q, nthreads=nthreads)) if orc else None
    report("k_fused_mf (strict)", "MFReconstruct::runReconstruction - IO", ms, B * (2 * P * 14 + P * 13), c, B * P, "pixels")

    out1 = eng._outputs(1, want_k=False)
    mf1 = mf[:1].contiguous()
    ms1 = timed(lambda: eng.run_mf(mf1, out=out1))
    rows_before = len(rows)
    report("k_fused_mf (strict), ONE scan: latency", "same, single scan (launch + 7 rows per CTA)", ms1, 2 * P * 14 + P * 13, c, P, "pixels")
    rows[rows_before]["speedup_vs_cpu_port"] = (c / (ms1 * 1e-3)) if c else None

    ms = timed(lambda: eng.run_mf(mf, mode=slr_b200.MODE_CORRECTED, out=out))
    c = cpu_time(lambda: orc.run_mf(h_mf[0], cams, Q, mode=1, nthreads=nthreads)) if orc else None
    report("k_fused_mf (corrected)", "(no reference counterpart)", ms, B * (2 * P * 14 + P * 13), c, B * P, "pixels")

    # ---------------- rectification ----------------
    import cv2
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    m1s, m2s = zip(*[cv2.convertMaps(xs * 1.004 - 3 + 2 * np.sin(ys / 90), ys * 0.998 + 1 + cam, cv2.CV_16SC2) for cam in range(2)])
    eng.set_rectify_maps(np.stack(m1s), np.stack(m2s))
    ms = timed(lambda: eng.rectify_stack(mf))
    c = cpu_time(lambda: [orc.remap_linear(h_mf[0, cam, n], m1s[cam], m2s[cam]) for cam in range(2) for n in range(14)]) if orc else None
    report("k0_rectify (14 planes)", "stereoRect::doStereoRectify (cv::remap)", ms, B * 2 * P * (14 * 2 + 6), c, B * 2 * 14 * P, "pixels")
    # raw stacks -> XYZ in one kernel (N1 as written): rectification inside the fused kernel's stage fill
    ms = timed(lambda: eng.run_mf_raw(mf, out=out))
    report("k_fused_flow<RAW> (rectify + strict MF pipeline, one kernel)", "doStereoRectify x 28 + MFReconstruct::runReconstruction - IO", ms,
           B * (2 * P * 14 + P * 13) + 2 * P * 6, None, B * P, "pixels")
    rect_tmp = torch.empty_like(mf)
    ms = timed(lambda: eng.run_mf(eng.rectify_stack(mf, out=rect_tmp), out=out))
    report("k0_rectify then k_fused_flow (two kernels)", "same", ms, B * (2 * P * 14 + P * 13) + 2 * P * 6, None, B * P, "pixels")
    del rect_tmp
    # ---------------- N4: PNG unfilter (Sub rows) + PointCloudImage layout; N2: merge ----------------
    filt = torch.randint(0, 256, (28, H, W + 1), dtype=torch.uint8, device="cuda")
    filt[:, :, 0] = 1
    planes = torch.empty((28, H, W), dtype=torch.uint8, device="cuda")
    lib = slr_b200.capi()
    import ctypes as C
    def unfilter_all():
        for k in range(28):
            lib.slr_png_unfilter(eng.h, C.c_void_p(filt[k].data_ptr()), C.c_void_p(planes[k].data_ptr()), 0)
    eng._bind_stream()
    ms = timed(unfilter_all)
    report("k_png_rows (28 images of Sub-filtered scanlines = one scan)", "PNG unfiltering inside cv::imread", ms, 28 * (P + H + P), None, 28 * P, "pixels")
    pts, src = eng.merge_scans(out[0], out[1])
    ms = timed(lambda: eng.merge_scans(out[0], out[1]))
    report("slr_merge_scans (ordered compaction + 3x4 transform)", "transfer_mat application + concatenation of the scans' clouds", ms,
           B * P * 13 + pts.shape[0] * 20, None, B * P, "pixels")
    # ---------------- K5: mesh indexing of one full-frame cloud (N3) ----------------
    import oracle_lib as _ol
    _o = orc if orc else _ol.load()
    sums_h, cnt_h = _o.pointcloud_from_dense(out[0][0].cpu().numpy(), out[1][0].cpu().numpy(), W, H)   # F7 drop rule applied
    sums_d, cnt_d = torch.from_numpy(sums_h).cuda(), torch.from_numpy(cnt_h).cuda()
    vert, src, faces = eng.mesh_index(sums_d, cnt_d, 0)
    nv, nf = vert.shape[0], faces.shape[0]
    ms = timed(lambda: eng.mesh_index(sums_d, cnt_d, 0))
    c = cpu_time(lambda: orc.mesh_index(sums_h, cnt_h, W, H, 0)) if orc else None
    rows_before = len(rows)
    report("k5 mesh index (1 cloud, 1280x1024)", "MeshCreator::exportPlyMesh index passes", ms, P * 13 + nv * 16 + nf * 12, c, P, "pixels")
    rows[rows_before]["speedup_vs_cpu_port"] = (c / (ms * 1e-3)) if c else None
    del mf, ph, mk, out

    # ---------------- 1280x1024 Gray EPI ----------------
    nb = slr_b200.gray_num_bits(W)
    h_g = np.stack([synth.synth_gray(W, H, seed=11 + s, integer_disparity=False, noise_dn=1.0) for s in range(min(B, 2))])
    g = torch.from_numpy(h_g).cuda().repeat((B + 1) // 2, 1, 1, 1, 1)[:B].contiguous()
    N = 2 + 2 * nb
    ms = timed(lambda: eng.gray_decode(g, nb, 0, 40, 3, W, H))
    c = cpu_time(lambda: [orc.gray_decode(h_g[0, cam], nb, 0, 40, 3, W, H) for cam in range(2)]) if orc else None
    report("k2_gray_decode (EPI, 11 bits)", "computeShadows+decodePatterns_GE+grayToDec", ms, B * 2 * P * (N + 5), c, B * 2 * P, "pixels")
    col, _, gm = eng.gray_decode(g, nb, 0, 40, 3, W, H)
    ms = timed(lambda: eng.match_triangulate_code(col, gm, want_k=False))
    col_h, gm_h = col[0].cpu().numpy(), gm[0].cpu().numpy()
    c = cpu_time(lambda: orc.ge_triangulate(col_h[0], gm_h[0], col_h[1], gm_h[1], Q, nthreads=nthreads)) if orc else None
    report("k3b code match+triangulate", "Reconstruct::triangulation_ge", ms, B * (2 * P * 5 + P * 13), c, B * P, "pixels")
    ms = timed(lambda: eng.run_ge(g, nb, 40, 3, W, want_k=False))
    report("slr_run_ge (k2 + k3b)", "Reconstruct::runReconstruction_GE - IO", ms, B * (2 * P * N + P * 13), None, B * P, "pixels")
    del g, col, gm
    eng.close()

    # ---------------- BASELINE configs 4 and 5: larger frames through slr_run_mf ----------------
    for (Wc, Hc, Fc, Sc, Bc, mode, label) in [(2048, 1536, 3, 4, 4, slr_b200.MODE_STRICT, "config 4: 2048x1536, 3x4, strict, one CTA per SM"),
                                               (4096, 3000, 4, 8, 2, slr_b200.MODE_CORRECTED, "config 5: 4096x3000, 4x8, corrected, K1 + plain K3a")]:
        ec = slr_b200.Engine(Wc, Hc, max_batch=Bc)
        camsc, Qc = slr_b200.synthetic_rig(Wc, Hc)
        ec.set_calib(camsc, Qc)
        Nc = 2 + Fc * Sc
        if Fc == 3 and Sc == 4:
            stc = ec.synth_mf(Bc, seed=3, integer_disparity=True, noise_dn=0.0)
        else:   # no reference pattern set exists for 4x8: smooth synthetic fringes generated on the device
            xs_ = torch.arange(Wc, device="cuda", dtype=torch.float32)[None, None, :] / Wc
            stc = torch.empty((Bc, 2, Nc, Hc, Wc), dtype=torch.uint8, device="cuda")
            stc[:, :, 0] = 220
            stc[:, :, 1] = 10
            for f_, fr in enumerate([70, 64, 59, 56][:Fc]):   # differences 6,5,3 -> 1,2 -> one beat period over the row
                for s_ in range(Sc):
                    for cam in range(2):
                        v = 128 + 90 * torch.cos(2 * np.pi * fr * (xs_ + 0.013 - 0.01 * cam) + 2 * np.pi * s_ / Sc)
                        stc[:, cam, 2 + Sc * f_ + s_] = v.to(torch.uint8)
        outc = ec._outputs(Bc, want_k=False)
        ms = timed(lambda: ec.run_mf(stc, F=Fc, S=Sc, mode=mode, out=outc))
        rows_before = len(rows)
        report(f"slr_run_mf ({label}, {Bc} scans)", "MFReconstruct::runReconstruction - IO", ms,
               Bc * (2 * Wc * Hc * Nc + Wc * Hc * 13), None, Bc * Wc * Hc, "pixels")
        del stc, outc
        ec.close()

    # ---------------- config 1: 640x480 Gray decode only ----------------
    W1, H1 = 640, 480
    e1 = slr_b200.Engine(W1, H1, max_batch=B)
    nb1 = slr_b200.gray_num_bits(W1)
    h_g1 = synth.synth_gray(W1, H1, seed=21, noise_dn=1.0)[None]
    g1 = torch.from_numpy(h_g1).cuda().repeat(B, 1, 1, 1, 1).contiguous()
    ms = timed(lambda: e1.gray_decode(g1, nb1, 0, 40, 3, W1, H1))
    c = cpu_time(lambda: [orc.gray_decode(h_g1[0, cam], nb1, 0, 40, 3, W1, H1) for cam in range(2)]) if orc else None
    report("k2_gray_decode (config 1: 640x480, 10 bits)", "same, BASELINE config 1", ms, B * 2 * W1 * H1 * (2 + 2 * nb1 + 5), c, B * 2 * W1 * H1, "pixels")
    e1.close()

    # ---------------- Gray-only bucket triangulation (small: latency bound) ----------------
    W2, H2 = 320, 240
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    e2 = slr_b200.Engine(W2, H2, max_batch=1)
    cams2 = cases.gray_only_rig(W2, H2)
    e2.set_calib(cams2, Q)
    h_go = synth.synth_gray(W2, H2, seed=31, rows=True, integer_disparity=True, noise_dn=1.0)[None]
    go = torch.from_numpy(h_go).cuda()
    nc2, nr2 = slr_b200.gray_num_bits(W2), slr_b200.gray_num_bits(H2)
    col, row, m = e2.gray_decode(go, nc2, nr2, 40, 3, W2, H2)
    ms = timed(lambda: e2.bucket_triangulate(col, row, m, W2, H2))
    d = [orc.gray_decode(h_go[0, cam], nc2, nr2, 40, 3, W2, H2) for cam in range(2)] if orc else None
    c = cpu_time(lambda: orc.gray_triangulate(d[0][0], d[0][1], d[0][2], d[1][0], d[1][1], d[1][2], W2, H2, cams2)) if orc else None
    rows_before = len(rows)
    report("k3c bucket triangulate (320x240, 1 scan)", "decodePaterns buckets + Reconstruct::triangulation", ms,
           2 * W2 * H2 * 9 + W2 * H2 * 13, c, W2 * H2, "cells")
    rows[rows_before]["speedup_vs_cpu_port"] = (c / (ms * 1e-3)) if c else None
    e2.close()

    # the same at the camera's full size (one scan)
    W3, H3 = 1280, 1024
    e3 = slr_b200.Engine(W3, H3, max_batch=1)
    cams3 = cases.gray_only_rig(W3, H3)
    e3.set_calib(cams3, Q)
    h_go3 = synth.synth_gray(W3, H3, seed=31, rows=True, integer_disparity=True, noise_dn=1.0)[None]
    go3 = torch.from_numpy(h_go3).cuda()
    nc3, nr3 = slr_b200.gray_num_bits(W3), slr_b200.gray_num_bits(H3)
    col3, row3, m3 = e3.gray_decode(go3, nc3, nr3, 40, 3, W3, H3)
    ms = timed(lambda: e3.bucket_triangulate(col3, row3, m3, W3, H3))
    d3 = [orc.gray_decode(h_go3[0, cam], nc3, nr3, 40, 3, W3, H3) for cam in range(2)] if orc else None
    c = cpu_time(lambda: orc.gray_triangulate(d3[0][0], d3[0][1], d3[0][2], d3[1][0], d3[1][1], d3[1][2], W3, H3, cams3)) if orc else None
    rows_before = len(rows)
    report("k3c bucket triangulate (1280x1024, 1 scan)", "decodePaterns buckets + Reconstruct::triangulation", ms,
           2 * W3 * H3 * 9 + W3 * H3 * 13, c, W3 * H3, "cells")
    rows[rows_before]["speedup_vs_cpu_port"] = (c / (ms * 1e-3)) if c else None
    e3.close()

    print(f"\n| kernel | replaces | ms ({B} scans) | algorithmic MB | GB/s | of {peak:.0f} GB/s ({peak_src.split(' ')[0]}) | CPU port s/scan ({nthreads} thr) | x CPU port |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        cpu = f"{r['cpu_port_s_per_scan']:.4f}" if r["cpu_port_s_per_scan"] else "-"
        sp = f"{r['speedup_vs_cpu_port']:.0f}" if r["speedup_vs_cpu_port"] else "-"
        print(f"| {r['kernel']} | {r['replaces']} | {r['ms']:.3f} | {r['algorithmic_MB']:.1f} | {r['GB/s']:.0f} | {100 * r['frac_of_hbm']:.1f} % | {cpu} | {sp} |")


if __name__ == "__main__":
    main()
