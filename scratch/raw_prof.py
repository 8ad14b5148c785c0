"""one launch of k_fused_flow<RAW> for ncu: python scratch/raw_prof.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, slr_b200, bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W, H = 1280, 1024
eng = slr_b200.Engine(W, H, max_batch=B, device=0)
cams, Q = slr_b200.synthetic_rig(W, H)
eng.set_calib(cams, Q)
m1, m2 = bench.rectify_maps_numpy(W, H)
eng.set_rectify_maps(m1, m2)
stack = eng.synth_mf(B, seed=1, integer_disparity=True, noise_dn=0.0)
out = eng._outputs(B, want_k=False)
for _ in range(3):
    eng.run_mf_raw(stack, 3, 4, 40, slr_b200.MODE_STRICT, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.run_mf_raw(stack, 3, 4, 40, slr_b200.MODE_STRICT, out=out)
b.record()
torch.cuda.synchronize()
print("ms", a.elapsed_time(b), "per scan", a.elapsed_time(b) / B)
