import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, slr_b200
Wc, Hc, Fc, Sc, Bc = 4096, 3000, 4, 8, 2
ec = slr_b200.Engine(Wc, Hc, max_batch=Bc)
cams, Q = slr_b200.synthetic_rig(Wc, Hc)
ec.set_calib(cams, Q)
Nc = 2 + Fc * Sc
xs_ = torch.arange(Wc, device="cuda", dtype=torch.float32)[None, None, :] / Wc
st = torch.empty((Bc, 2, Nc, Hc, Wc), dtype=torch.uint8, device="cuda")
st[:, :, 0] = 220
st[:, :, 1] = 10
for f_, fr in enumerate([70, 64, 59, 56][:Fc]):
    for s_ in range(Sc):
        for cam in range(2):
            v = 128 + 90 * torch.cos(2 * np.pi * fr * (xs_ + 0.013 - 0.01 * cam) + 2 * np.pi * s_ / Sc)
            st[:, cam, 2 + Sc * f_ + s_] = v.to(torch.uint8)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ph, mk = ec.mf_decode(st, F=Fc, S=Sc, mode=1)
print("k1 corrected 4x8 ms:", timed(lambda: ec.mf_decode(st, F=Fc, S=Sc, mode=1)))
print("k3a plain ms:", timed(lambda: ec.match_triangulate_phase(ph, mk, want_k=False)))
p0 = ph[0, :, 100].cpu().numpy(); m0 = mk[0, :, 100].cpu().numpy()
print("valid L/R", m0.sum(1), "distinct R", len(np.unique(p0[1][m0[1] > 0])), "range", np.nanmin(p0[1]), np.nanmax(p0[1]))
srt = np.sort(p0[1][m0[1] > 0]); pl = p0[0][m0[0] > 0]
nm = np.searchsorted(srt, pl + 0.1) - np.searchsorted(srt, pl - 0.1)
print("true matches per left px: mean %.1f max %d" % (nm.mean(), nm.max()))
