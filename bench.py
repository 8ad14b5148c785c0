#!/usr/bin/env python
"""bench.py — Mpoints/s triangulated on synthetic 1280x1024 stereo 3-freq x 4-shift stacks (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--config 3|4|5] [--impl reference]

One "step" = the multi-frequency pipeline (shadow mask + strict-mode phase decode + heterodyne + per-row phase
correspondence + Q-matrix triangulation; Duke/mfreconstruct.cpp:160-334) over one batch of B synthetic scans per GPU.
Rank 0 prints ONE JSON line:

  value           whole-job Mpoints/s, inputs resident in HBM, device-timed (CUDA events), max over ranks
  roofline        algorithmic bytes of the step / event-timed kernel duration vs the measured HBM copy peak
  config.variants the same step on the unfriendly inputs (sensor noise sigma 2 DN + sub-pixel disparity; corrected mode)
  e2e             same metric through the C-ABI host-buffer call at the SAME batch (pinned host stacks in, clouds out,
                  copies inside the timed region), with its own roofline against the pinned-copy rate measured here
  cpu_baseline    N = 1: the reference's own code (oracle/_ref, Duke/*.cpp compiled unmodified) on this box's host cores
  with_allgather  N > 1: the step with the cloud of every rank assembled on every rank — by the kernel's own peer
                  stores over NVLink (no collective call), and by one NCCL all-gather for comparison
  row_bands       N > 1: ONE scan split into N row bands, single-scan latency

`--config 4|5` selects BASELINE.json's larger configurations (on-device synthesis): 2048x1536 3x4, 16 scans over the
ranks; 4096x3000 4x8 (corrected mode: the reference hard-codes 3x4), 64 scans over the ranks.
`--impl reference` times the reference's CPU implementation alone (rank 0; the other ranks exit).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLACK_THR = 40
METRIC = "Mpoints/s triangulated (1280x1024 stereo, 3-freq x 4-shift)"
UNIT = "Mpoints/s"

# BASELINE.json configs.  3 = the configuration the metric is quoted on (configs[1..2] at 1280x1024); 4, 5 = configs[3..4].
CONFIGS = {
    3: dict(W=1280, H=1024, F=3, S=4, strict=True, total_scans=None,
            workload="MF pipeline (strict decode + phase match + Q triangulation), 1280x1024 stereo 3x4"),
    4: dict(W=2048, H=1536, F=3, S=4, strict=True, total_scans=16,
            workload="BASELINE config 4: MF pipeline, 2048x1536 stereo 3x4, batch of 16 scans tiled across the GPUs"),
    5: dict(W=4096, H=3000, F=4, S=8, strict=False, total_scans=64,
            workload="BASELINE config 5: MF pipeline, 4096x3000 stereo 4x8 (corrected mode), batch of 64 scans sharded "
                     "across the GPUs"),
}


def algorithmic_bytes_per_scan(cfg):
    """every input byte once + every final output byte once = C*P*N stack bytes + P*(12 xyz + 1 valid)."""
    P = cfg["W"] * cfg["H"]
    return 2 * P * (2 + cfg["F"] * cfg["S"]) + P * 13


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(cfg_id, scans_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json).
    Only a capture taken at THIS configuration and batch counts; anything else reports null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        if int(d.get("config", 3)) == cfg_id and int(d["scans_per_launch"]) == scans_per_launch:
            return int(d["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def bind_to_gpu_numa_node(index):
    """N > 1: run this rank's host threads (and first-touch its pinned staging buffers) on the CPUs NVML reports as
    local to the GPU.  Returns the number of CPUs bound to, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class stdout_to_stderr:
    """NCCL prints its version banner on stdout when a communicator is created; the bench's stdout is ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the scheduler, not OpenMP)."""
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def make_host_inputs(n_scans, rank, noise_dn=0.0, integer_disparity=True):
    """The 1280x1024 3x4 workload, identical for both arms: numpy scene synthesis (slr_b200.synth), seeds by rank and
    scan.  Returns uint8 [n_scans, 2, 14, H, W]."""
    import numpy as np
    from slr_b200 import synth
    c = CONFIGS[3]
    return np.stack([synth.synth_mf(c["W"], c["H"], seed=1000 * rank + 1 + s, integer_disparity=integer_disparity,
                                    noise_dn=noise_dn) for s in range(n_scans)])


def rectify_maps_numpy(W, H):
    """CV_16SC2 maps as cv::convertMaps / cv::initUndistortRectifyMap produce them (integer source pixel + 5+5 fractional
    bits), for a small rotation + scale + lens-like bow of each camera: int16 [2][H][W][2], uint16 [2][H][W]."""
    import numpy as np
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    m1, m2 = [], []
    for cam in range(2):
        th = (0.4 - 0.8 * cam) * np.pi / 180
        cx, cy = W / 2, H / 2
        dx, dy = xs - cx, ys - cy
        r2 = (dx * dx + dy * dy) / (cx * cx + cy * cy)
        sc = 1.004 + 0.01 * r2
        mx = cx + sc * (np.cos(th) * dx - np.sin(th) * dy) + 1.5
        my = cy + sc * (np.sin(th) * dx + np.cos(th) * dy) - 0.75
        ix, iy = np.rint(mx * 32).astype(np.int64), np.rint(my * 32).astype(np.int64)
        m1.append(np.stack([ix >> 5, iy >> 5], -1).astype(np.int16))
        m2.append(((iy & 31) * 32 + (ix & 31)).astype(np.uint16))
    return np.stack(m1), np.stack(m2)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms: the reference's own code (oracle/_ref/libref.so = Duke/*.cpp compiled unmodified against oracle/ref_shim)
# and the oracle port of the same loops.  Only bench legs that are reported as baselines call into oracle/.
# ---------------------------------------------------------------------------------------------------------------------
def _ref_band_worker(task):
    """One process = one single-threaded instance of the reference (it is not re-entrant: file-scope toggles and
    globals, Duke/mfreconstruct.cpp:4-5): computeShadows + decodePatterns of both cameras + triangulation of a band."""
    band, cams, Q = task
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_lib
    ref = ref_lib.load()
    dec = [ref.mf_decode(band[c], BLACK_THR) for c in range(2)]
    _, cnt = ref.mf_triangulate(dec[0][0], dec[0][1], dec[1][0], dec[1][1], cams, Q)
    return int(cnt.sum())


def reference_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref.so"))


class ReferencePool:
    """The real reference on `nproc` host cores: independent processes, each running row bands of the same scans."""

    def __init__(self, nproc):
        import multiprocessing as mp
        self.nproc = nproc
        self.pool = mp.get_context("fork").Pool(nproc) if nproc > 1 else None

    def run(self, stack, cams, Q, rows_per_band, bands):
        """`bands` bands of `rows_per_band` rows, taken round-robin from the scans of `stack`; returns (points, seconds)."""
        H = stack.shape[3]
        tasks = []
        for k in range(bands):
            s, r0 = k % stack.shape[0], (k * rows_per_band) % (H - rows_per_band + 1)
            tasks.append((stack[s, :, :, r0:r0 + rows_per_band].copy(), cams, Q))
        t0 = time.perf_counter()
        counts = self.pool.map(_ref_band_worker, tasks, chunksize=1) if self.pool else [_ref_band_worker(t) for t in tasks]
        return sum(counts), time.perf_counter() - t0

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def port_rate(stack, cams, Q, nthreads, min_seconds, max_scans):
    """Oracle port of the reference loops (same O(W^2) first-match search, rows over `nthreads` OpenMP threads)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    orc = oracle_lib.load()
    pts, scans, t0 = 0, 0, time.perf_counter()
    while True:
        pts += orc.run_mf(stack[scans % stack.shape[0]], cams, Q, black_thr=BLACK_THR, mode=0, nthreads=nthreads)[3]
        scans += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or scans >= max_scans:
            return pts / dt / 1e6, scans, dt


def cpu_baseline(stack, cams, Q):
    """cpu_baseline of the CUDA arm's line (N = 1, rank 0): ~20 s of CPU work on a bounded sample of the same scans."""
    nthr = host_threads()
    out = {"unit": UNIT}
    if reference_available():
        rows = 32
        pool = ReferencePool(nthr)
        pool.run(stack, cams, Q, 4, nthr)                                   # start the workers, load the library
        pts, dt = pool.run(stack, cams, Q, rows, 4 * nthr)
        pool.close()
        p1, d1 = ReferencePool(1).run(stack, cams, Q, rows, 4)
        out.update({"value": pts / dt / 1e6, "cores": nthr, "kind": "reference",
                    "sample": f"{4 * nthr} row bands of {rows} x 1280 pixels of the same scans in {dt:.1f} s: "
                              f"oracle/_ref (Duke/*.cpp compiled unmodified), one single-threaded process per core",
                    "single_thread_value": p1 / d1 / 1e6,
                    "single_thread_sample": f"4 bands of {rows} rows in {d1:.1f} s (the reference runs single-threaded)"})
    vp, sp, dp = port_rate(stack, cams, Q, nthr, 6.0, 2000)
    v1, s1, d1 = port_rate(stack, cams, Q, 1, 4.0, 50)
    port = {"value": vp, "cores": nthr, "kind": "port", "sample": f"{sp} scans in {dp:.1f} s, rows over {nthr} OpenMP threads",
            "single_thread_value": v1, "single_thread_sample": f"{s1} scans in {d1:.1f} s"}
    if "value" in out:
        out["port"] = port     # the C restatement: no per-pixel std::vector churn, hence ~15x the real reference
    else:
        out.update(port)
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, alone."""
    import slr_b200
    cfg = CONFIGS[3]
    cams, Q = slr_b200.synthetic_rig(cfg["W"], cfg["H"])
    stack = make_host_inputs(2, 0)          # the first scans of rank 0's batch in the CUDA arm
    nthr = host_threads()
    if reference_available():
        rows, kind = 16, "reference"
        pool = ReferencePool(nthr)
        step = lambda: pool.run(stack, cams, Q, rows, nthr)                 # noqa: E731
        sample = (f"{nthr} row bands of {rows} x 1280 pixels per step (one per core) of the same synthetic scans; "
                  "oracle/_ref = Duke/mfreconstruct.cpp etc. compiled unmodified, one single-threaded process per core")
    else:
        kind, pool = "port", None
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        orc = oracle_lib.load()

        def step():
            t0 = time.perf_counter()
            n = orc.run_mf(stack[0], cams, Q, black_thr=BLACK_THR, mode=0, nthreads=nthr)[3]
            return n, time.perf_counter() - t0
        sample = f"one scan per step, oracle port (oracle/_ref was not built here), rows over {nthr} OpenMP threads"
    for _ in range(args.warmup):
        step()
    pts, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        pts += step()[0]
    dt = time.perf_counter() - t0
    if pool:
        pool.close()
    val = pts / dt / 1e6
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def drop_in_call(stack_one_scan, W, H, calls=8):
    """The reference's own entry point on its own input format: a project directory (28 PNG files as cv::imwrite writes
    them + the calibration text files) through the drop-in MFReconstruct::runReconstruction (facade_demo drives it as
    MainWindow::startreconstruct does), in steady state: image files -> PointCloudImage, nothing skipped.
    Returns None when the facade is not built."""
    import re
    import shutil
    import tempfile
    import numpy as np
    from slr_b200 import synth
    demo = os.path.join(ROOT, "structure-light-reconstructor_b200", "facade", "facade_demo")
    if not os.path.exists(demo):
        return None
    tmp = tempfile.mkdtemp(prefix="slr_bench_")
    try:
        def write_mat(path, m):
            with open(path, "w") as f:
                for r in np.atleast_2d(np.asarray(m, np.float64)):
                    f.write("\t".join(f"{v:.9g}" for v in r) + "\t\n")
        for side, fx, cx, cy, dist, R, t in (
                ("left", 2400.0, W / 2 + 2.25, H / 2 - 1.5, [-0.11, 0.07, 0.0006, -0.0003, 0.0], np.eye(3), [0, 0, 0]),
                ("right", 2396.0, W / 2 - 1.75, H / 2 + 1.25, [-0.09, 0.04, -0.0004, 0.0005, 0.0],
                 [[0.9998, 0.0, 0.02], [0.0, 1.0, 0.0], [-0.02, 0.0, 0.9998]], [-200.0, 0.5, 1.0])):
            d = os.path.join(tmp, "calib", side)
            os.makedirs(d)
            K = [[fx, 0, cx], [0, fx + 1, cy], [0, 0, 1]]
            for name, m in (("cam_matrix.txt", K), ("cam_distortion.txt", np.array(dist)[:, None]), ("cam_rotation_matrix.txt", R),
                            ("cam_trans_vectror.txt", np.array(t)[:, None]), ("cam_stereo.txt", K),
                            ("distortion_stereo.txt", np.array(dist)[:, None])):
                write_mat(os.path.join(d, name), m)
        write_mat(os.path.join(tmp, "calib", "R_stereo.txt"), [[0.99985, 0.002, 0.0172], [-0.0021, 0.999995, 0.004], [-0.0172, -0.004, 0.99984]])
        write_mat(os.path.join(tmp, "calib", "T_stereo.txt"), np.array([-200.0, 0.4, 1.2])[:, None])
        for name in ("fundamental_stereo.txt", "H1_mat.txt", "H2_mat.txt"):
            write_mat(os.path.join(tmp, "calib", name), np.eye(3))
        png_bytes = 0
        for cam, side, pre in ((0, "left", "L"), (1, "right", "R")):
            d = os.path.join(tmp, "scan", side, "0")
            os.makedirs(d)
            for i in range(stack_one_scan.shape[1]):
                synth.write_png_opencv_style(os.path.join(d, f"{pre}{i}.png"), stack_one_scan[cam, i])
                png_bytes += os.path.getsize(os.path.join(d, f"{pre}{i}.png"))
        env = dict(os.environ, DUKE_REPEAT=str(calls + 2))
        r = subprocess.run([demo, "mf", tmp, "0", str(W), str(H), str(W), str(H), str(BLACK_THR), "0", "0",
                            os.path.join(tmp, "out.bin")], capture_output=True, text=True, env=env, timeout=300)
        ms = [float(x) for x in re.findall(r"runReconstruction #\d+ .*?: ([0-9.]+) ms", r.stderr)]
        pts = re.search(r"mf: (\d+) points", r.stdout)
        if r.returncode != 0 or len(ms) < 3 or not pts:
            return {"error": (r.stderr or r.stdout)[-300:]}
        steady = statistics.median(ms[2:])
        return {"ms_per_call": steady, "first_call_ms": ms[0], "points_per_call": int(pts.group(1)),
                "value": int(pts.group(1)) / (steady * 1e-3) / 1e6, "unit": UNIT, "png_bytes_per_scan": png_bytes,
                "host_threads": host_threads(),
                "what": "MFReconstruct::runReconstruction of the drop-in classes on a project directory: 28 PNG files (Sub filter, "
                        "Z_RLE, as cv::imwrite writes them) -> inflate on the host threads -> GPU (PNG unfilter, rectification "
                        "inside the fused kernel, decode, match, triangulation, PointCloudImage layout) -> PointCloudImage; median "
                        f"of calls 3..{len(ms)} of one process"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="scans per GPU per step (config 3 default: 64)")
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the assembly legs (peer stores, NCCL, row bands)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import slr_b200
    from slr_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()

    cfg = CONFIGS[args.config]
    W, H, F, S = cfg["W"], cfg["H"], cfg["F"], cfg["S"]
    n_img = 2 + F * S
    mode = slr_b200.MODE_STRICT if cfg["strict"] else slr_b200.MODE_CORRECTED
    if args.batch:
        B = args.batch
    elif cfg["total_scans"]:
        B = max(1, cfg["total_scans"] // world)
    else:
        B = 64
    dev = torch.device("cuda", local_rank)
    eng = slr_b200.Engine(W, H, max_batch=B, device=local_rank)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tile(t, n):
        return t if t.shape[0] >= n else t.repeat((n + t.shape[0] - 1) // t.shape[0], 1, 1, 1, 1)[:n].contiguous()

    # inputs resident in HBM before the timed region.  Config 3: 8 distinct host-synthesised scans (294 MB > L2) tiled
    # to B, shared with the CPU arms; configs 4, 5: synthesised on the device (up to 53 GB per GPU).
    h_in = None
    if args.config == 3:
        h_in = make_host_inputs(min(B, 8), rank)
        stack = tile(torch.from_numpy(h_in).cuda(), B)
    else:
        stack = eng.synth_mf(B, seed=1000 * rank + 1, integer_disparity=True, noise_dn=0.0, F=F, S=S)
    out = eng._outputs(B, want_k=False)
    torch.cuda.synchronize()

    def timed_steps(stk, steps, md, outputs):
        """W warm-ups, then `steps` steps bracketed by barrier + synchronize.
        -> (total ms, [per-step ms], points per step, kernels launched inside the timed region)"""
        for _ in range(args.warmup):
            outputs[4].zero_()
            eng.run_mf(stk, F, S, BLACK_THR, md, out=outputs)
        torch.cuda.synchronize()
        pts = int(outputs[4].item())
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        l0 = eng.launches()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for a, b in evs:
            a.record()
            eng.run_mf(stk, F, S, BLACK_THR, md, out=outputs)
            b.record()
        t1.record()
        barrier()
        return t0.elapsed_time(t1), [a.elapsed_time(b) for a, b in evs], pts, eng.launches() - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, step_ms, points_per_step, launches = timed_steps(stack, args.steps, mode, out)
    clocks = sampler.stop() if rank == 0 else None
    alg = algorithmic_bytes_per_scan(cfg) * B
    peak, peak_src = load_peak()

    # ---- the same step on unfriendly inputs (N = 1, config 3) ----
    variants = []
    if world == 1 and args.config == 3 and not args.no_variants:
        v_steps = max(3, min(args.steps, 10))
        noisy = tile(torch.from_numpy(make_host_inputs(min(B, 8), rank, noise_dn=2.0, integer_disparity=False)).cuda(), B)
        for name, stk, md in (("sensor noise sigma 2 DN + sub-pixel disparity, strict mode", noisy, mode),
                              ("corrected mode (atan2 + heterodyne cascade; no reference counterpart), noise-free", stack,
                               slr_b200.MODE_CORRECTED)):
            ms, per, pts, _ = timed_steps(stk, v_steps, md, out)
            variants.append({"name": name, "value": pts / (ms / v_steps * 1e-3) / 1e6, "unit": UNIT,
                             "ms_per_step": ms / v_steps, "points_per_step": pts,
                             "roofline_frac": alg / (statistics.mean(per) * 1e-3) / 1e9 / peak})
        del noisy
        # RAW camera stacks: rectification inside the fused kernel's stage fill (slr_run_mf_raw, one
        # kernel) against K0 into HBM followed by the fused kernel (two kernels, the stack written and read once more)
        m1, m2 = rectify_maps_numpy(W, H)
        eng.set_rectify_maps(m1, m2)
        raw_leg = {}
        for name, fn in (("fused", lambda: eng.run_mf_raw(stack, F, S, BLACK_THR, mode, out=out)),
                         ("k0_then_fused", lambda: eng.run_mf(eng.rectify_stack(stack, out=rect_buf), F, S, BLACK_THR, mode, out=out))):
            if name == "k0_then_fused":
                rect_buf = torch.empty_like(stack)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            l0 = eng.launches()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(v_steps):
                out[4].zero_()
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / v_steps
            raw_leg[name] = {"ms_per_step": ms, "value": int(out[4].item()) / (ms * 1e-3) / 1e6, "unit": UNIT,
                             "kernels_per_step": (eng.launches() - l0) / v_steps,
                             "roofline_frac": (alg + 2 * W * H * 6) / (ms * 1e-3) / 1e9 / peak}
        del rect_buf
        # single-scan latency (one scan is launch- and tail-bound, far below the L2 size)
        one = eng._outputs(1, want_k=False)
        ms, per, _, _ = timed_steps(stack[:1], 20, mode, one)
        single_scan_ms = statistics.median(per)
    else:
        single_scan_ms, raw_leg = None, None

    # ---- N > 1: assemble every rank's cloud on every rank ----
    gather = {}
    if world > 1 and not args.no_gather:
        recv = (world - 1) * B * H * W * 13
        g_steps = max(3, min(args.steps, 10))
        n_pts = torch.zeros(1, dtype=torch.int64, device=dev)
        # (a) no collective call: peer-mapped buffers, the fused kernel's epilogue stores every row to all ranks
        asm = parallel.PeerAssembly(eng, B, H, W, slots=2)
        nothing = (None, None, None, None, n_pts)
        for slot in range(2):
            asm.select(slot)
            eng.run_mf(stack, F, S, BLACK_THR, mode, out=nothing)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for k in range(g_steps):
            asm.select(k & 1)
            eng.run_mf(stack, F, S, BLACK_THR, mode, out=nothing)
        g1.record()
        barrier()
        gather["peer_ms"] = g0.elapsed_time(g1) / g_steps
        xa, va = asm.views((g_steps - 1) & 1)
        gather["peer_points_seen"] = int(va.sum().item())       # the whole job's cloud is on this rank
        asm.close()
        # (b) one NCCL all-gather per step (xyz and valid of a slot in one group call), on a side stream
        ids = [slr_b200.Engine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        with stdout_to_stderr():
            comm = eng.nccl_comm_create(world, rank, ids[0])
        casm = parallel.CloudAssembly(B, H, W, dev, slots=2, native=(eng, comm))
        outs = []
        for slot in range(2):
            xl, vl = casm.local_views(slot)
            outs.append((xl, vl, None, None, n_pts))
        for slot in range(2):
            eng.run_mf(stack, F, S, BLACK_THR, mode, out=outs[slot])
            casm.gather(slot)
        casm.wait()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for k in range(g_steps):
            casm.before_compute(k & 1)
            eng.run_mf(stack, F, S, BLACK_THR, mode, out=outs[k & 1])
            casm.gather(k & 1)
        casm.wait()
        g1.record()
        barrier()
        gather["nccl_ms"] = g0.elapsed_time(g1) / g_steps
        gather["recv_bytes"] = recv
        del casm, outs
        eng.nccl_comm_destroy(comm)
        # (c) ONE scan split into `world` row bands (both cameras' rows lo..hi-1 per rank), assembled by peer stores
        if H % world == 0 and args.config != 5:
            hb = H // world
            band = slr_b200.Engine(W, hb, max_batch=1, device=local_rank)
            band.set_calib(cams, Q)
            band.set_row_offset(rank * hb)
            basm = parallel.PeerAssembly(band, 1, hb, W, slots=2)
            bstack = stack[:1, :, :, rank * hb:(rank + 1) * hb].contiguous()
            for slot in range(2):
                basm.select(slot)
                band.run_mf(bstack, F, S, BLACK_THR, mode, out=nothing)
            lat = []
            for k in range(20):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                basm.select(k & 1)
                band.run_mf(bstack, F, S, BLACK_THR, mode, out=nothing)
                b.record()
                torch.cuda.synchronize()
                lat.append(a.elapsed_time(b))
            gather["band_ms"] = statistics.median(lat)
            gather["band_rows"] = hb
            basm.close()
            band.close()

    # ---- end to end through the host-buffer C-ABI call, same batch ----
    e2e = None
    if not args.no_e2e:
        Be = B
        h_stack = torch.empty((Be, 2, n_img, H, W), dtype=torch.uint8).pin_memory()
        h_stack.copy_(stack[:Be])
        h_xyz = torch.empty((Be, H, W, 3), dtype=torch.float32).pin_memory()
        h_valid = torch.empty((Be, H, W), dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()
        # the copy rates this host gives this rank (all ranks copy at once, as in the timed leg): pinned H2D and D2H
        # in parallel on two streams, which is what the pipeline does
        probe_d = torch.empty_like(stack[:min(Be, 8)])
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            probe_d.copy_(h_stack[:probe_d.shape[0]], non_blocking=True)
        with torch.cuda.stream(s_out):
            h_xyz[:probe_d.shape[0]].copy_(out[0][:probe_d.shape[0]], non_blocking=True)
        torch.cuda.synchronize()
        probe_s = time.perf_counter() - t0
        h2d_rate = probe_d.numel() / probe_s / 1e9
        del probe_d
        e_steps = max(3, min(args.steps, 5))
        eng.run_mf_host(h_stack, h_xyz, h_valid, None, F, S, BLACK_THR, mode)
        barrier()
        t0 = time.perf_counter()
        pts = 0
        for _ in range(e_steps):
            pts += eng.run_mf_host(h_stack, h_xyz, h_valid, None, F, S, BLACK_THR, mode)
        torch.cuda.synchronize()
        e2e = {"pts": pts, "dt": time.perf_counter() - t0, "steps": e_steps, "batch": Be,
               "h2d": Be * 2 * n_img * H * W, "d2h": Be * H * W * 13, "h2d_rate": h2d_rate}
        del h_stack, h_xyz, h_valid
        if world == 1 and args.config == 3 and rank == 0:
            e2e["drop_in"] = drop_in_call(make_host_inputs(1, rank, noise_dn=2.0, integer_disparity=False)[0], W, H)

    # ---- reduce over ranks: time = max, points = sum ----
    g_keys = [k for k in ("peer_ms", "nccl_ms", "band_ms") if k in gather]
    if world > 1:
        t = torch.tensor([total_ms, e2e["dt"] if e2e else 0.0] + [gather[k] for k in g_keys], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p = torch.tensor([points_per_step, e2e["pts"] if e2e else 0], device="cuda", dtype=torch.int64)
        dist.all_reduce(p, op=dist.ReduceOp.SUM)
        r = torch.tensor([e2e["h2d_rate"] if e2e else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(r, op=dist.ReduceOp.MIN)
        total_ms, e_dt = float(t[0]), float(t[1])
        for i, k in enumerate(g_keys):
            gather[k] = float(t[2 + i])
        points_all, e_pts_all, h2d_rate = int(p[0]), int(p[1]), float(r[0])
    else:
        e_dt = e2e["dt"] if e2e else 0.0
        points_all, e_pts_all = points_per_step, (e2e["pts"] if e2e else 0)
        h2d_rate = e2e["h2d_rate"] if e2e else 0.0

    if rank == 0:
        ms_per_step = total_ms / args.steps
        achieved = alg / (statistics.mean(step_ms) * 1e-3) / 1e9
        config = {"workload": cfg["workload"], "scans_per_gpu_per_step": B,
                  "mode": "strict (reference arithmetic)" if cfg["strict"] else "corrected (oracle-only: no reference counterpart)",
                  "inputs": "noise-free, integer disparity (the friendliest case; see variants)",
                  "black_threshold": BLACK_THR,
                  "l2": f"inputs {B * 2 * n_img * W * H / 1e6:.0f} MB per step exceed the 126 MB L2",
                  "parallelism": f"scans sharded over {world} GPU(s); no collective inside 'value' (see with_allgather)",
                  "host_cpus_bound_per_rank": numa, "points_per_step": points_all,
                  "mpixels_per_s": world * B * W * H / (ms_per_step * 1e-3) / 1e6}
        if variants:
            config["variants"] = variants
        if single_scan_ms is not None:
            config["single_scan_latency_ms"] = single_scan_ms
            config["raw_input"] = dict(raw_leg, note="the same stacks taken as RAW camera images: stereoRect::doStereoRectify "
                                       "(cv::remap, CV_16SC2 maps of a 0.4 degree / 0.4 % warp) inside the fused kernel's "
                                       "stage fill (one launch, the rectified stack never exists in HBM) vs as a separate "
                                       "pass (K0 writes it, the fused kernel reads it back); algorithmic bytes = the step's + 6 map bytes per "
                                       "pixel and camera")
        result = {
            "metric": METRIC, "value": points_all / (ms_per_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(args.config, B), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "kernel_ms_mean": statistics.mean(step_ms),
                         "kernel_ms_median": statistics.median(step_ms), "kernels_per_step": launches / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e:
            ms_e = e_dt / e2e["steps"] * 1e3
            result["e2e"] = {"value": e_pts_all / e_dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                             "d2h_bytes_per_step": e2e["d2h"], "scans_per_step": e2e["batch"], "steps": e2e["steps"],
                             "ms_per_step": ms_e,
                             "roofline": {"bound": "pcie (pinned host -> device copy of the stacks; the cloud returns on the "
                                                   "other direction concurrently)",
                                          "achieved": e2e["h2d"] / (ms_e * 1e-3) / 1e9, "peak": h2d_rate, "unit": "GB/s",
                                          "frac": e2e["h2d"] / (ms_e * 1e-3) / 1e9 / h2d_rate if h2d_rate else None,
                                          "peak_source": "pinned copy of the same buffers measured in this process, both "
                                                         "directions at once, all ranks at once (min over ranks)"}}
            if e2e.get("drop_in"):
                result["e2e"]["drop_in_call"] = e2e["drop_in"]
        if gather:
            w = {"bytes_received_per_rank_per_step": gather["recv_bytes"]}
            for key, name, note in (("peer_ms", "peer_stores", "no collective call: every rank's k_fused_flow stores each output row "
                                     "into all ranks' peer-mapped clouds over NVLink from its epilogue"),
                                    ("nccl_ms", "nccl", "one slr_allgather (ncclAllGather of xyz + valid in one group) per step on a "
                                     "side stream, double buffered")):
                if key in gather:
                    w[name] = {"value": points_all / (gather[key] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": gather[key],
                               "nvlink_recv_gbs_per_gpu": gather["recv_bytes"] / (gather[key] * 1e-3) / 1e9, "note": note}
            w["points_seen_on_rank0"] = gather.get("peer_points_seen")
            result["with_allgather"] = w
            if "band_ms" in gather:
                result["row_bands"] = {"single_scan_latency_ms": gather["band_ms"], "rows_per_gpu": gather["band_rows"],
                                       "note": "ONE scan split into row bands over the GPUs, assembled on every GPU by peer stores"}
        if world == 1 and args.config == 3 and not args.no_cpu:
            result["cpu_baseline"] = cpu_baseline(h_in[:2], cams, Q)
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
