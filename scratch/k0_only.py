import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, cv2, slr_b200
from slr_b200 import synth
W, H, B = 1280, 1024, 8
eng = slr_b200.Engine(W, H, max_batch=B)
xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
m1s, m2s = zip(*[cv2.convertMaps(xs * 1.004 - 3 + 2 * np.sin(ys / 90), ys * 0.998 + 1 + cam, cv2.CV_16SC2) for cam in range(2)])
eng.set_rectify_maps(np.stack(m1s), np.stack(m2s))
mf = torch.randint(0, 255, (B, 2, 14, H, W), dtype=torch.uint8, device="cuda")
for _ in range(3):
    out = eng.rectify_stack(mf)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): eng.rectify_stack(mf)
b.record(); torch.cuda.synchronize()
print("k0 ms", a.elapsed_time(b) / 5)
