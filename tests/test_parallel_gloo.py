"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2: scan sharding and the all-gather
that assembles the per-rank clouds (the N > 1 path of bench.py uses the same helpers over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slr_b200 import parallel


def test_scan_shard_partitions_the_batch():
    for n in (0, 1, 7, 16, 64):
        for world in (1, 2, 3, 4, 8):
            blocks = [parallel.scan_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.scan_shard(4, 2, 2)


def test_row_band_partitions_one_scan():
    for h in (1, 48, 1024, 1536, 3000):
        for world in (1, 2, 4, 8):
            bands = [parallel.row_band(h, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_scans, H, W, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.scan_shard(n_scans, rank, world)
        g = torch.Generator().manual_seed(1234)
        full_xyz = torch.randn((n_scans, H, W, 3), generator=g)
        full_valid = (torch.rand((n_scans, H, W), generator=g) > 0.3).to(torch.uint8)
        xyz, valid = parallel.gather_clouds(full_xyz[lo:hi].clone(), full_valid[lo:hi].clone(), n_scans)
        ok = torch.equal(xyz, full_xyz) and torch.equal(valid, full_valid)
        # weak-scaling bookkeeping as bench.py does it: time = max over ranks, points = sum over ranks
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p = torch.tensor([int(full_valid[lo:hi].sum())], dtype=torch.int64)
        dist.all_reduce(p, op=dist.ReduceOp.SUM)
        ok = ok and float(t) == float(world) and int(p) == int(full_valid.sum())
        if n_scans % world == 0:   # equal shards: the pre-allocated in-place assembly bench.py --gather uses
            b = n_scans // world
            asm = parallel.CloudAssembly(b, H, W, "cpu", slots=2)
            for slot in range(2):
                xl, vl = asm.local_views(slot)
                xl.copy_(full_xyz[lo:hi])
                vl.copy_(full_valid[lo:hi])
                gx, gv = asm.gather(slot)
                ok = ok and torch.equal(gx, full_xyz) and torch.equal(gv, full_valid)
            asm.wait()
        open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_scans", [4, 5])
def test_gather_clouds_world2_gloo(tmp_path, n_scans):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_scans, 6, 16, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]
