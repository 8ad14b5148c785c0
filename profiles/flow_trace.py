"""Job timeline of k_fused_flow from the clock64 stamps of the trace build (make -C csrc TRACE=1).
Run on the GPU box:  python profiles/flow_trace.py   -> gpurun_out/flow_trace.bin + a summary on stdout."""
import os, sys, struct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SLR_B200_LIB", os.path.join(ROOT, "structure-light-reconstructor_b200", "libslr_b200_trc.so"))
out = os.path.join(ROOT, "gpurun_out", "flow_trace.bin")


def capture():
    import torch
    import slr_b200
    W, H, B = 1280, 1024, 16
    eng = slr_b200.Engine(W, H, max_batch=B, device=0)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    stack = eng.synth_mf(B, seed=1, integer_disparity=True, noise_dn=0.0)
    for _ in range(3):
        eng.run_mf(stack, black_thr=40)
    torch.cuda.synchronize()
    os.environ["SLR_FLOW_TRACE_OUT"] = out
    eng.run_mf(stack, black_thr=40)
    torch.cuda.synchronize()


def report(path):
    raw = open(path, "rb").read()
    n, n_d, n_q, warps = struct.unpack("4i", raw[:16])
    d = np.frombuffer(raw[16:], np.int64).reshape(n, 4)
    d = d[d[:, 1] != 0]
    typ = d[:, 0] & 0xff
    row = (d[:, 0] >> 8) & 0xffffffff
    t0 = d[:, 1].min()
    draw, ready, end = (d[:, k] - t0 for k in (1, 2, 3))
    rows = int(row.max()) + 1
    step = (end.max() - draw.min()) / rows
    print(f"jobs {len(d)}  n_d {n_d} n_q {n_q} warps {warps}  rows {rows}  cycles/row {step:.0f}")
    for name, sel in (("decode", typ == 0), ("query", typ >= 1)):
        w, run = (ready - draw)[sel], (end - ready)[sel]
        print(f"{name:7s} jobs {sel.sum():5d}  wait mean {w.mean():7.0f} p50 {np.median(w):6.0f} p90 {np.percentile(w, 90):6.0f} max {w.max():6d}"
              f"   run mean {run.mean():7.0f} p90 {np.percentile(run, 90):6.0f}   share of warp time waiting {w.sum() / (w.sum() + run.sum()):.3f}")
    # where a warp's time goes: running a job, waiting for its dependencies, or between jobs (draw + loop overhead)
    warp = (d[:, 0] >> 40) & 0xff
    tot_run = (end - ready).sum()
    tot_wait = (ready - draw).sum()
    span = 0
    for w in np.unique(warp):
        sel = warp == w
        span += end[sel].max() - draw[sel].min()
    print(f"warp time: run {tot_run / span:.3f}  wait {tot_wait / span:.3f}  between jobs {(span - tot_run - tot_wait) / span:.3f}"
          f"   (decode run {((end - ready)[typ == 0]).sum() / span:.3f}, query run {((end - ready)[typ >= 1]).sum() / span:.3f})")
    print("per-row timeline (cycles from the row's first decode draw), rows 20..25:")
    for r in range(20, min(26, rows)):
        dsel, qsel = (typ == 0) & (row == r), (typ >= 1) & (row == r)
        if not dsel.any() or not qsel.any():
            continue
        base = draw[dsel].min()
        print(f"  row {r}: D draw {draw[dsel].min() - base:6d}..{draw[dsel].max() - base:6d} ready {ready[dsel].min() - base:6d}..{ready[dsel].max() - base:6d} "
              f"end ..{end[dsel].max() - base:6d} | Q draw {draw[qsel].min() - base:6d}..{draw[qsel].max() - base:6d} ready {ready[qsel].min() - base:6d} end ..{end[qsel].max() - base:6d}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        report(sys.argv[1])
    else:
        capture()
        if os.path.exists(out):
            report(out)
        else:
            print("no trace (not the TRACE=1 build)")
