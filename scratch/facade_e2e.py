"""Drop-in end to end at full size: a 1280x1024 project directory (28 PNGs + calibration text files) through
facade_demo = MainWindow::startreconstruct minus the GUI, incl. the PLY export.  Prints facade_demo's timings."""
import os, sys, tempfile, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_gpu_facade as tf
from slr_b200 import synth
W, H = 1280, 1024
tmp = tempfile.mkdtemp()
stacks = synth.synth_mf(W, H, seed=5, noise_dn=1.0)
tf.make_project(tmp, W, H, 0, stacks)
for rep in range(1):
    t0 = time.perf_counter()
    env = dict(os.environ, DUKE_EXPORT_PLY=os.path.join(tmp, "cloud.ply"), DUKE_TIMING="1", DUKE_REPEAT="3")
    r = subprocess.run([tf.DEMO, "mf", tmp, "0", str(W), str(H), str(W), str(H), "40", "0", "0", os.path.join(tmp, "out.bin")],
                       capture_output=True, text=True, env=env)
    print(f"run {rep}: process wall {1e3 * (time.perf_counter() - t0):.0f} ms\n" + r.stderr.strip() + "\n" + r.stdout.strip())
print("ply bytes", os.path.getsize(os.path.join(tmp, "cloud.ply")))
