"""Multi-GPU plumbing (host side only).  Scans are independent, so ranks own contiguous blocks of the batch and
no collective is needed on the data path; the optional all-gather assembles every rank's cloud on every rank
(`torch.distributed.all_gather_into_tensor`: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def scan_shard(n_scans: int, rank: int, world: int):
    """Contiguous block [lo, hi) of the batch owned by `rank` (B/G scans per rank; the first
    n_scans % world ranks take one extra)."""
    if world <= 0 or not (0 <= rank < world) or n_scans < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_scans, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_clouds(xyz_local, valid_local, n_scans: int, group=None):
    """All-gather per-rank clouds [b_local, H, W, 3] / [b_local, H, W] into [n_scans, ...] on every rank.
    Ranks may own different scan counts (ragged shards are padded to the largest shard for the collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [scan_shard(n_scans, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    H, W = valid_local.shape[1:]

    def padded(t, tail):
        out = t.new_zeros((bmax,) + tail)
        out[: t.shape[0]] = t
        return out
    xl, vl = padded(xyz_local, (H, W, 3)), padded(valid_local, (H, W))
    xg = xl.new_empty((world * bmax, H, W, 3))
    vg = vl.new_empty((world * bmax, H, W))
    dist.all_gather_into_tensor(xg, xl, group=group)
    dist.all_gather_into_tensor(vg, vl, group=group)
    keep = torch.cat([torch.arange(r * bmax, r * bmax + (hi - lo)) for r, (lo, hi) in enumerate(sizes)]).to(xg.device)
    return xg.index_select(0, keep), vg.index_select(0, keep)
