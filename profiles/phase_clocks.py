"""Phase timeline of k_fused_mf from the clock64 stamps of the debug build (make -C csrc PHASE=1).
Run on the GPU box:  SLR_B200_LIB=.../libslr_b200_dbg.so python scratch/phase_clocks.py [strict|corrected]"""
import os, sys, struct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SLR_B200_LIB", os.path.join(ROOT, "structure-light-reconstructor_b200", "libslr_b200_dbg.so"))
import torch
import slr_b200
mode = 1 if (len(sys.argv) > 1 and sys.argv[1] == "corrected") else 0
W, H, B = 1280, 1024, 16
eng = slr_b200.Engine(W, H, max_batch=B, device=0)
cams, Q = slr_b200.synthetic_rig(W, H)
eng.set_calib(cams, Q)
stack = eng.synth_mf(B, seed=1, integer_disparity=True, noise_dn=0.0)
for _ in range(3):
    eng.run_mf(stack, black_thr=40, mode=mode)
torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", f"phase_clocks_{'corrected' if mode else 'strict'}.bin")
os.environ["SLR_PHASE_CLOCKS_OUT"] = out
eng.run_mf(stack, black_thr=40, mode=mode)
torch.cuda.synchronize()
if not os.path.exists(out):
    print("no stamps (not the debug build)"); sys.exit(0)
raw = open(out, "rb").read()
grid, rows, warps, pts = struct.unpack("4i", raw[:16])
d = np.frombuffer(raw[16:], np.int64).reshape(grid, rows, warps, pts)
nw = int((d[0, 0, :, 0] != 0).sum())
d = d[:, :, :nw, :].astype(np.float64)
print(f"grid {grid} warps/CTA {nw} rows sampled {rows}")
# per CTA-row: barrier release times = max over warps of the pre-barrier stamp (approx. = min of the post stamp)
t_start = d[..., 1].min(-1)            # decode starts (after barrier 1)
t_dec_end_w = d[..., 2]                # each warp's decode end
t_bar2 = d[..., 3].min(-1)
t_q_end_w = d[..., 4]
t_bar3 = d[..., 5].min(-1)
t_clear0 = d[..., 0]                   # after clear, before mbar wait
row_time = np.diff(d[..., 5].min(-1), axis=1)
print("cycles per row per CTA: mean %.0f" % row_time.mean())
dec = t_bar2 - t_start
qry = t_bar3 - t_bar2
clr = (t_start[:, 1:] - t_bar3[:, :-1])
print("decode phase %.0f  query phase %.0f  clear+tma wait %.0f" % (dec.mean(), qry.mean(), clr.mean()))
print("decode: warp busy mean %.0f (first-done %.0f, last-done %.0f)" % ((t_dec_end_w - t_start[..., None]).mean(),
      (t_dec_end_w - t_start[..., None]).min(-1).mean(), (t_dec_end_w - t_start[..., None]).max(-1).mean()))
print("query : warp busy mean %.0f (first-done %.0f, last-done %.0f)" % ((t_q_end_w - t_bar2[..., None]).mean(),
      (t_q_end_w - t_bar2[..., None]).min(-1).mean(), (t_q_end_w - t_bar2[..., None]).max(-1).mean()))
print("clear : after-clear stamp - bar3 mean %.0f ; mbar wait+bar1 %.0f" % ((t_clear0[:, 1:] - t_bar3[:, :-1, None]).mean(),
      (t_start[:, 1:, None] - t_clear0[:, 1:]).mean()))

# are the two CTAs of an SM in lock step?  overlap of their decode intervals
smid = d[:, 0, 0, 6].astype(int)
from collections import defaultdict
by = defaultdict(list)
for c in range(grid):
    by[smid[c]].append(c)
ov = []
for sm, cs in by.items():
    if len(cs) != 2:
        continue
    a, b = cs
    # decode intervals of a, b over the sampled rows
    for r in range(rows):
        a0, a1 = t_start[a, r], t_bar2[a, r]
        best = 0.0
        for r2 in range(rows):
            b0, b1 = t_start[b, r2], t_bar2[b, r2]
            best = max(best, max(0.0, min(a1, b1) - max(a0, b0)))
        ov.append(best / (a1 - a0))
print("SMs with 2 CTAs: %d; mean fraction of a CTA's decode phase overlapped by the co-resident CTA's decode phase: %.2f"
      % (sum(1 for v in by.values() if len(v) == 2), float(np.mean(ov)) if ov else -1))
print("(anti-phase = ~0, independent = ~%.2f, lock step = ~1)" % (dec.mean() / row_time.mean()))
