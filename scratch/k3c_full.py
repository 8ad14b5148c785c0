import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, slr_b200, cases
from slr_b200 import synth
for (W, H) in [(320, 240), (1280, 1024)]:
    e = slr_b200.Engine(W, H, max_batch=1)
    cams = cases.gray_only_rig(W, H)
    _, Q = slr_b200.synthetic_rig(W, H)
    e.set_calib(cams, Q)
    g = torch.from_numpy(synth.synth_gray(W, H, seed=31, rows=True, integer_disparity=True, noise_dn=1.0)[None]).cuda()
    nc, nr = slr_b200.gray_num_bits(W), slr_b200.gray_num_bits(H)
    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    col, row, m = e.gray_decode(g, nc, nr, 40, 3, W, H)
    print(W, H, "k2 (col+row) ms", timed(lambda: e.gray_decode(g, nc, nr, 40, 3, W, H)), "k3c ms", timed(lambda: e.bucket_triangulate(col, row, m, W, H)))
    s, c, n = e.bucket_triangulate(col, row, m, W, H)
    print("  cells", int(n.item()), "launches per call", None)
    e.close()
