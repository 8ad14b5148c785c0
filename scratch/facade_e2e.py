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
stacks = synth.synth_mf(W, H, seed=5, integer_disparity=False, noise_dn=2.0)
tf.make_project(tmp, W, H, 0, stacks)
# the scan images as OpenCV's encoder writes them (cv::imwrite: Sub filter on every row, Z_RLE, 8 KB IDAT chunks)
rng = np.random.default_rng(0)
for cam, side, pre in ((0, "left", "L"), (1, "right", "R")):
    for i in range(14):
        tf._write_png_filtered(os.path.join(tmp, "scan", side, "0", f"{pre}{i}.png"), stacks[cam, i], "sub", rng)
print("png bytes per image:", [os.path.getsize(os.path.join(tmp, "scan", "left", "0", f"L{i}.png")) for i in (0, 2, 7)])
for rep, host in enumerate((False, True)):
    t0 = time.perf_counter()
    env = dict(os.environ, DUKE_EXPORT_PLY=os.path.join(tmp, "cloud.ply"), DUKE_TIMING="1", DUKE_REPEAT="4")
    if host:
        env["DUKE_HOST_DECODE"] = "1"      # round-1 route: host decode of every image, slr_run_mf_host, addDense
    r = subprocess.run([tf.DEMO, "mf", tmp, "0", str(W), str(H), str(W), str(H), "40", "0", "0", os.path.join(tmp, "out.bin")],
                       capture_output=True, text=True, env=env)
    print(f"run {rep} ({'host decode' if host else 'GPU ingest'}): process wall {1e3 * (time.perf_counter() - t0):.0f} ms\n" + r.stderr.strip() + "\n" + r.stdout.strip())
print("ply bytes", os.path.getsize(os.path.join(tmp, "cloud.ply")))
