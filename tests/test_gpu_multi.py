"""Two-GPU test of the one collective of the path, through the C ABI: slr_allgather (ncclAllGather bound at run time by
libslr_b200.so) assembles the per-rank clouds in place.  Skipped on single-GPU boxes (the gloo tests cover the host
logic there); run with `gpurun --gpus 2`."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, B = 320, 24, 2


def _worker(rank, world, tmp):
    import torch
    import slr_b200
    torch.cuda.set_device(rank)
    eng = slr_b200.Engine(W, H, max_batch=B, device=rank)
    cams, Q = slr_b200.synthetic_rig(W, H)
    eng.set_calib(cams, Q)
    idf = os.path.join(tmp, "nccl_id.bin")
    if rank == 0:
        uid = slr_b200.Engine.nccl_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            assert time.time() - t0 < 60
            time.sleep(0.05)
        uid = open(idf, "rb").read()
    comm = eng.nccl_comm_create(world, rank, uid)
    xyz_all = torch.full((world * B, H, W, 3), 7.0, dtype=torch.float32, device="cuda")
    valid_all = torch.full((world * B, H, W), 9, dtype=torch.uint8, device="cuda")
    stacks = [eng.synth_mf(B, seed=100 + r, integer_disparity=True, noise_dn=1.0) for r in range(world)]
    n = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = (xyz_all[rank * B:(rank + 1) * B], valid_all[rank * B:(rank + 1) * B], None, None, n)
    eng.run_mf(stacks[rank], black_thr=40, want_k=False, out=out)       # the kernel writes straight into block `rank`
    eng.allgather(comm, world, rank, B, xyz_all, valid_all)
    torch.cuda.synchronize()
    ok = True
    for r in range(world):                                               # every block == that rank's scans recomputed here
        xyz, valid, _, _ = eng.run_mf(stacks[r], black_thr=40)
        torch.cuda.synchronize()
        ok = ok and torch.equal(valid_all[r * B:(r + 1) * B], valid)
        a = xyz_all[r * B:(r + 1) * B].cpu().numpy().view(np.uint32)
        ok = ok and bool((a == xyz.cpu().numpy().view(np.uint32)).all())
    eng.nccl_comm_destroy(comm)
    eng.close()
    open(os.path.join(tmp, f"rank{rank}.{'ok' if ok else 'bad'}"), "w").close()


def test_slr_allgather_assembles_the_clouds_of_two_gpus(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    mp.spawn(_worker, args=(2, str(tmp_path)), nprocs=2, join=True)
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("rank")) == ["rank0.ok", "rank1.ok"]
