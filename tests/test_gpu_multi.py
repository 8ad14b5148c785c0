"""Two-GPU tests of the cloud assembly, through the C ABI: (1) slr_allgather (ncclAllGather bound at run time by
libslr_b200.so) assembles the per-rank clouds in place; (2) the collective-free form: peer-mapped buffers registered as
gather targets, the fused kernel's epilogue stores every row to both GPUs; (3) one scan split into two row bands.
Skipped on single-GPU boxes (the gloo tests cover the host logic there); run with `gpurun --gpus 2`."""
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


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _peer_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    import slr_b200
    from slr_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # carries the IPC handles and the barriers only
    ok = True
    try:
        # ---- (2) scans sharded over the ranks, assembled by the kernel's peer stores ----
        eng = slr_b200.Engine(W, H, max_batch=B, device=rank)
        cams, Q = slr_b200.synthetic_rig(W, H)
        eng.set_calib(cams, Q)
        stacks = [eng.synth_mf(B, seed=200 + r, integer_disparity=False, noise_dn=1.0) for r in range(world)]
        asm = parallel.PeerAssembly(eng, B, H, W, slots=2)
        for slot in range(2):
            xa, va = asm.views(slot)
            xa.fill_(7.0)
            va.fill_(9)
        torch.cuda.synchronize()
        dist.barrier()
        n = torch.zeros(1, dtype=torch.int64, device="cuda")
        for slot in range(2):
            asm.select(slot)
            eng.run_mf(stacks[rank], black_thr=40, want_k=False, out=(None, None, None, None, n))
        torch.cuda.synchronize()
        dist.barrier()
        eng.set_gather_targets([], [])
        for slot in range(2):
            xa, va = asm.views(slot)
            for r in range(world):     # every block == that rank's scans recomputed here
                xyz, valid, _, _ = eng.run_mf(stacks[r], black_thr=40)
                torch.cuda.synchronize()
                ok = ok and torch.equal(va[r * B:(r + 1) * B], valid)
                ok = ok and bool((xa[r * B:(r + 1) * B].cpu().numpy().view(np.uint32) == xyz.cpu().numpy().view(np.uint32)).all())
        asm.close()
        eng.close()

        # ---- (3) ONE scan split into row bands (MF and Gray-EPI), assembled the same way ----
        Hf = 2 * H
        full = slr_b200.Engine(W, Hf, max_batch=1, device=rank)
        cams, Q = slr_b200.synthetic_rig(W, Hf)
        full.set_calib(cams, Q)
        lo, hi = parallel.row_band(Hf, rank, world)
        band = slr_b200.Engine(W, hi - lo, max_batch=1, device=rank)
        band.set_calib(cams, Q)
        band.set_row_offset(lo)
        asm = parallel.PeerAssembly(band, 1, hi - lo, W, slots=1)
        asm.select(0)
        mf = full.synth_mf(1, seed=300, integer_disparity=False, noise_dn=1.0)
        band.run_mf(mf[:, :, :, lo:hi].contiguous(), black_thr=40, want_k=False, out=(None, None, None, None, n))
        torch.cuda.synchronize()
        dist.barrier()
        xyz, valid, _, _ = full.run_mf(mf, black_thr=40)
        xa, va = asm.views(0)
        ok = ok and torch.equal(va.reshape(1, Hf, W), valid)
        ok = ok and bool((xa.reshape(1, Hf, W, 3).cpu().numpy().view(np.uint32) == xyz.cpu().numpy().view(np.uint32)).all())
        dist.barrier()
        nb = slr_b200.gray_num_bits(W)
        ge = full.synth_gray(1, seed=301, integer_disparity=False, noise_dn=2.0)
        band.run_ge(ge[:, :, :, lo:hi].contiguous(), nb, black_thr=40, white_thr=3, want_k=False,
                    out=(None, None, None, None, n))
        torch.cuda.synchronize()
        dist.barrier()
        xyz, valid, _, _, _ = full.run_ge(ge, nb, black_thr=40, white_thr=3)
        ok = ok and torch.equal(va.reshape(1, Hf, W), valid)
        ok = ok and bool((xa.reshape(1, Hf, W, 3).cpu().numpy().view(np.uint32) == xyz.cpu().numpy().view(np.uint32)).all())
        asm.close()
        band.close()
        full.close()
    finally:
        dist.destroy_process_group()
    open(os.path.join(tmp, f"rank{rank}.{'ok' if ok else 'bad'}"), "w").close()


def test_peer_store_assembly_and_row_bands_on_two_gpus(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    mp.spawn(_peer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("rank")) == ["rank0.ok", "rank1.ok"]
