"""one launch each of the round-2 secondary kernels, for ncu: K0 (warp job), PNG unfilter, PointCloudImage layout, merge"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, slr_b200, bench
B, W, H = 8, 1280, 1024
eng = slr_b200.Engine(W, H, max_batch=B, device=0)
cams, Q = slr_b200.synthetic_rig(W, H)
eng.set_calib(cams, Q)
m1, m2 = bench.rectify_maps_numpy(W, H)
eng.set_rectify_maps(m1, m2)
stack = eng.synth_mf(B, seed=1, integer_disparity=False, noise_dn=2.0)
for _ in range(2):
    rect = eng.rectify_stack(stack)
    xyz, valid, k, n = eng.run_mf(rect, want_k=False)
    pts, src = eng.merge_scans(xyz, valid)
    one = stack[0].cpu().numpy()
    imgs = [np.concatenate([np.ones((H, 1), np.uint8), np.diff(one[c, i].astype(np.int16), axis=1, prepend=0).astype(np.uint8)], 1)
            for c in range(2) for i in range(14)]
    h_sum = np.empty((H, W, 3), np.float32); h_cnt = np.empty((H, W), np.uint8)
    eng.run_mf_ingested(imgs, [True] * 28, h_sum, h_cnt, scan_w=W, scan_h=H)
torch.cuda.synchronize()
print("ok", pts.shape)
