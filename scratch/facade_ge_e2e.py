"""Drop-in Gray-EPI call at full size: 1280x1024, 2 x 24 PNGs (OpenCV-style) through facade_demo ge."""
import os, sys, tempfile, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_gpu_facade as tf
from slr_b200 import synth
W, H = 1280, 1024
tmp = tempfile.mkdtemp()
stacks = synth.synth_gray(W, H, seed=5, integer_disparity=False, noise_dn=2.0)
tf.make_project(tmp, W, H, 0, stacks[:, :2])
for cam, side, pre in ((0, "left", "L"), (1, "right", "R")):
    for i in range(stacks.shape[1]):
        synth.write_png_opencv_style(os.path.join(tmp, "scan", side, "0", f"{pre}{i}.png"), stacks[cam, i])
env = dict(os.environ, DUKE_TIMING="1", DUKE_REPEAT="4")
r = subprocess.run([tf.DEMO, "ge", tmp, "0", str(W), str(H), str(W), str(H), "40", "4", "1", os.path.join(tmp, "out.bin")],
                   capture_output=True, text=True, env=env)
print(r.stderr.strip()[-3000:], r.stdout.strip())
