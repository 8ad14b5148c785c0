#!/usr/bin/env python
"""bench.py — Mpoints/s triangulated on synthetic 1280x1024 stereo 3-freq x 4-shift stacks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = the multi-frequency pipeline (shadow mask + strict-mode phase decode + heterodyne +
per-row phase correspondence + Q-matrix triangulation; Duke/mfreconstruct.cpp:160-334) over one
batch of B synthetic scans per GPU.  Prints ONE JSON line (rank 0).

  value        whole-job Mpoints/s, inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e          same metric through the C-ABI host-buffer call (pinned host stacks in, cloud out;
               host<->device copies inside the timed region)
  roofline     algorithmic bytes of the step / event-timed duration vs the measured HBM copy peak
  cpu_baseline the oracle port of the reference loop on this box's host cores (bounded sample)

`--impl reference` times the reference's own CPU algorithm (the oracle port: same O(W^2) first-match
search, rows spread over all host threads) on a bounded sample of the same workload.
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

W, H, F, S = 1280, 1024, 3, 4
N_IMG = 2 + F * S
BLACK_THR = 40
METRIC = "Mpoints/s triangulated (1280x1024 stereo, 3-freq x 4-shift)"
UNIT = "Mpoints/s"


def algorithmic_bytes_per_scan():
    # every input byte once + every final output byte once:
    # C*P*N stack bytes + P*(12 xyz + 1 valid)
    P = W * H
    return 2 * P * N_IMG + P * 13


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(scans_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), scaled
    from the capture's scans per launch to this run's."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return int(d["dram_bytes_per_launch"] * scans_per_launch / d.get("scans_per_launch", 16))
        except Exception:
            return None
    return None


def bind_to_gpu_numa_node(index):
    """N > 1: run this rank's host threads (and first-touch its pinned staging buffers) on the CPUs NVML reports as
    local to the GPU, so that the end-to-end leg's host<->device copies do not cross the socket interconnect.
    Returns the number of CPUs bound to, or None when NVML / the affinity call is unavailable."""
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
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(n_scans, rank):
    """The synthetic workload, identical for both arms: numpy scene synthesis (slr_b200.synth), seeds by
    rank and scan.  Returns uint8 [n_scans, 2, 14, H, W]."""
    import numpy as np
    from slr_b200 import synth
    return np.stack([synth.synth_mf(W, H, seed=1000 * rank + 1 + s, integer_disparity=True, noise_dn=0.0)
                     for s in range(n_scans)])


def cpu_port_rate(stack_np, cams, Q, nthreads, min_seconds, max_scans):
    """Oracle port of the reference loop on `stack_np` [B,2,14,H,W]; returns (Mpoints/s, scans, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    orc = oracle_lib.load()
    if not nthreads:
        nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    pts, scans, t0 = 0, 0, time.perf_counter()
    while True:
        _, _, _, n = orc.run_mf(stack_np[scans % stack_np.shape[0]], cams, Q, F=F, S=S, black_thr=BLACK_THR, mode=0,
                                nthreads=nthreads)
        pts += n
        scans += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or scans >= max_scans:
            break
    return pts / dt / 1e6, scans, dt, nthreads


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port, all host threads)."""
    if rank != 0:
        return
    import numpy as np
    import slr_b200
    from slr_b200 import synth
    cams, Q = slr_b200.synthetic_rig(W, H)
    nscan = 2
    stack = make_inputs(nscan, 0)   # the first scans of rank 0's batch in the CUDA arm
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    orc = oracle_lib.load()
    # all host threads this process may use; torchrun exports OMP_NUM_THREADS=1, so ask the scheduler, not OpenMP
    nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = 2     # scans per step (bounded sample of the B-scan workload)
    for _ in range(args.warmup):
        orc.run_mf(stack[0], cams, Q, nthreads=nthreads)
    pts = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        for s in range(per_step):
            pts += orc.run_mf(stack[(k * per_step + s) % nscan], cams, Q, nthreads=nthreads)[3]
    dt = time.perf_counter() - t0
    val = pts / dt / 1e6
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MF pipeline (strict decode + phase match + Q triangulation), 1280x1024 stereo 3x4",
                   "scans_per_step": per_step, "note": "the reference cannot be compiled here (Qt5/OpenCV 2.4.9/"
                   "windows.h); this is the oracle port of Duke/mfreconstruct.cpp:160-334, rows over all host threads; "
                   "it omits the reference's per-pixel std::vector churn, so it is faster than the real reference"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": f"{per_step} scans/step x {args.steps} steps of the same synthetic workload"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="scans per GPU per step (64: launch + tail amortised; 16 costs 3 %%)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", action="store_true",
                    help="N > 1: also time the steps followed by an NCCL all-gather of every rank's cloud (reported "
                         "under 'with_allgather'; never part of 'value')")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import slr_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = args.batch
    eng = slr_b200.Engine(W, H, max_batch=B, device=local_rank)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    # inputs resident in HBM before the timed region; B*36.7 MB exceeds the 126 MB L2 for B >= 4
    n_distinct = min(B, 8)      # distinct scans (8 x 36.7 MB = 294 MB > L2), tiled to B
    h_in = make_inputs(n_distinct, rank)
    stack = torch.from_numpy(h_in).cuda()
    if B > n_distinct:
        stack = stack.repeat((B + n_distinct - 1) // n_distinct, 1, 1, 1, 1)[:B].contiguous()
    out = eng._outputs(B, want_k=False)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out[4].zero_()
        eng.run_mf(stack, F, S, BLACK_THR, slr_b200.MODE_STRICT, out=out)
    torch.cuda.synchronize()
    points_per_step = int(out[4].item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = eng.launches()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        evs[k][0].record()
        eng.run_mf(stack, F, S, BLACK_THR, slr_b200.MODE_STRICT, out=out)
        evs[k][1].record()
    t_end.record()
    barrier()
    launches = eng.launches() - l0
    total_ms = t_start.elapsed_time(t_end)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    clocks = sampler.stop() if rank == 0 else None

    # ---- optional: the same steps followed by an all-gather of the cloud over NVLink (north star's assembly) ----
    gather = None
    if args.gather and world > 1:
        from slr_b200 import parallel
        # the collective goes through the C ABI (slr_allgather: NCCL bound by the library); torch.distributed only
        # carries the 128-byte NCCL id to the other ranks
        ids = [slr_b200.Engine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = eng.nccl_comm_create(world, rank, ids[0])
        asm = parallel.CloudAssembly(B, H, W, torch.device("cuda", local_rank), slots=2, native=(eng, comm))
        outs = []
        for slot in range(2):   # the kernels write straight into this rank's block of the assembled cloud
            xl, vl = asm.local_views(slot)
            outs.append((xl, vl, out[2], out[3], out[4]))
        for slot in range(2):   # warm-up: NCCL channel setup stays outside the timed region
            eng.run_mf(stack, F, S, BLACK_THR, slr_b200.MODE_STRICT, out=outs[slot])
            asm.gather(slot)
        asm.wait()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for k in range(args.steps):
            slot = k & 1
            asm.before_compute(slot)
            eng.run_mf(stack, F, S, BLACK_THR, slr_b200.MODE_STRICT, out=outs[slot])
            asm.gather(slot)    # in place, on the communication stream: overlaps the next step's kernel
        asm.wait()
        g1.record()
        barrier()
        gather = {"ms": g0.elapsed_time(g1), "bytes_received_per_rank_per_step": (world - 1) * B * H * W * 13}

    # ---- end to end through the host-buffer C-ABI call ----
    e2e = None
    if not args.no_e2e:
        Be = min(B, 8)
        h_stack = torch.empty((Be, 2, N_IMG, H, W), dtype=torch.uint8).pin_memory()
        h_stack.copy_(stack[:Be])
        h_xyz = torch.empty((Be, H, W, 3), dtype=torch.float32).pin_memory()
        h_valid = torch.empty((Be, H, W), dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()
        e_steps = max(3, min(args.steps, 10))
        eng.run_mf_host(h_stack, h_xyz, h_valid, None, F, S, BLACK_THR, slr_b200.MODE_STRICT)
        barrier()
        t0 = time.perf_counter()
        pts = 0
        for _ in range(e_steps):
            pts += eng.run_mf_host(h_stack, h_xyz, h_valid, None, F, S, BLACK_THR, slr_b200.MODE_STRICT)
        torch.cuda.synchronize()
        e_dt = time.perf_counter() - t0
        e2e = {"pts": pts, "dt": e_dt, "steps": e_steps, "batch": Be,
               "h2d": Be * 2 * N_IMG * H * W, "d2h": Be * H * W * 13}

    # ---- reduce over ranks: time = max, points = sum ----
    if world > 1:
        t = torch.tensor([total_ms, e2e["dt"] if e2e else 0.0, gather["ms"] if gather else 0.0], device="cuda",
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p = torch.tensor([points_per_step, e2e["pts"] if e2e else 0], device="cuda", dtype=torch.int64)
        dist.all_reduce(p, op=dist.ReduceOp.SUM)
        total_ms, e_dt_max = float(t[0]), float(t[1])
        if gather:
            gather["ms"] = float(t[2])
        points_all, e_pts_all = int(p[0]), int(p[1])
    else:
        e_dt_max = e2e["dt"] if e2e else 0.0
        points_all, e_pts_all = points_per_step, (e2e["pts"] if e2e else 0)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = points_all / (ms_per_step * 1e-3) / 1e6
        peak, peak_src = load_peak()
        alg = algorithmic_bytes_per_scan() * B
        med_ms = statistics.median(step_ms)
        achieved = alg / (statistics.mean(step_ms) * 1e-3) / 1e9
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MF pipeline (strict decode + phase match + Q triangulation), 1280x1024 stereo 3x4",
                       "scans_per_gpu_per_step": B, "mode": "strict (reference arithmetic)", "black_threshold": BLACK_THR,
                       "l2": f"inputs {B * 2 * N_IMG * W * H / 1e6:.0f} MB per step exceed the 126 MB L2",
                       "parallelism": f"scans sharded over {world} GPU(s), no data-path collective",
                       "host_cpus_bound_per_rank": numa,
                       "points_per_step": points_all, "mpixels_per_s": world * B * W * H / (ms_per_step * 1e-3) / 1e6},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(B), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "kernel_ms_mean": statistics.mean(step_ms),
                         "kernel_ms_median": med_ms,
                         "kernels_per_step": launches / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if e2e:
            result["e2e"] = {"value": e_pts_all / e_dt_max / 1e6, "unit": UNIT,
                             "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                             "scans_per_step": e2e["batch"], "steps": e2e["steps"],
                             "ms_per_step": e_dt_max / e2e["steps"] * 1e3}
        if gather:
            g_ms = gather["ms"] / args.steps
            result["with_allgather"] = {"value": points_all / (g_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": g_ms,
                                        "bytes_received_per_rank_per_step": gather["bytes_received_per_rank_per_step"],
                                        "note": "every step followed by slr_allgather (in-place ncclAllGather of xyz+valid over NVLink through the C ABI) on a side stream, double buffered; not in 'value'"}
        if world == 1 and not args.no_cpu:
            h = h_in[:2]
            v1, sc1, dt1, nt1 = cpu_port_rate(h, cams, Q, 0, 10.0, 2000)
            vs, scs, dts, _ = cpu_port_rate(h, cams, Q, 1, 5.0, 50)
            result["cpu_baseline"] = {"value": v1, "unit": UNIT, "cores": nt1, "kind": "port",
                                      "sample": f"{sc1} scans of the same workload in {dt1:.1f} s, rows over {nt1} threads",
                                      "single_thread_value": vs,
                                      "single_thread_sample": f"{scs} scans in {dts:.1f} s (the reference is single-threaded)"}
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
