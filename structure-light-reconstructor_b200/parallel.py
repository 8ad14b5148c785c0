"""Multi-GPU plumbing (host side only).  Scans are independent, so ranks own contiguous blocks of the batch (or, for
one scan, contiguous row bands) and no collective is needed on the data path.  Assembling every rank's cloud on every
rank is done either by `PeerAssembly` (the kernels store their output rows straight into every peer's buffer over
NVLink: no collective call at all) or by `CloudAssembly` / `gather_clouds` (an all-gather: NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations


def scan_shard(n_scans: int, rank: int, world: int):
    """Contiguous block [lo, hi) of the batch owned by `rank` (B/G scans per rank; the first
    n_scans % world ranks take one extra)."""
    if world <= 0 or not (0 <= rank < world) or n_scans < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_scans, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def row_band(height: int, rank: int, world: int):
    """Rows [lo, hi) of ONE scan owned by `rank`: both cameras' rows lo..hi-1 (the match runs along rectified rows, so
    bands need no halo; SURVEY.md §8e unit 2).  Same contiguous split as scan_shard."""
    return scan_shard(height, rank, world)


class _DevicePtr:
    """Zero-copy torch view of raw device memory through the CUDA array interface."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerAssembly:
    """The assembled cloud of all ranks WITHOUT a collective: every rank allocates [world * b_local, H, W, 3] xyz +
    [world * b_local, H, W] valid with `slr_peer_alloc`, maps every other rank's buffers through their CUDA IPC
    handles, and registers all of them as gather targets of its engine.  `slr_run_mf` then stores each output row
    into block `rank` of EVERY rank's buffer from the kernel's epilogue (peer stores over NVLink / NVSwitch), so the
    transfer overlaps the kernel that produces the data.  After the kernels of all ranks have completed and the
    ranks have synchronised (`dist.barrier()`), `views(slot)` holds the whole cloud on every rank.
    For row bands, pass the band height as H: block r is then band r of the scan and the assembled tensor, viewed as
    [b_local, world * H, W, ...] for b_local == 1, is the full image.  Equal shards only."""

    def __init__(self, eng, b_local: int, H: int, W: int, slots: int = 2, group=None):
        import torch
        import torch.distributed as dist
        self.eng, self.b, self.slots = eng, b_local, slots
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n = self.world * b_local
        px = n * H * W
        self.own, self.peers, self._views = [], [], []
        handles = []
        for _ in range(slots):
            xp, xh = eng.peer_alloc(px * 12)
            vp, vh = eng.peer_alloc(px)
            self.own.append((xp, vp))
            handles.append((xh, vh))
            self._views.append((torch.as_tensor(_DevicePtr(xp, (n, H, W, 3), "<f4"), device=eng.dev),
                                torch.as_tensor(_DevicePtr(vp, (n, H, W), "|u1"), device=eng.dev)))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=group)
        for slot in range(slots):
            mapped = []
            # rank r walks its targets in the order r+1, r+2, ...: at any moment every GPU is written by ONE peer (all
            # ranks starting with rank 0's buffer would queue seven writers on one NVLink ingress at a time)
            for k in range(1, self.world):
                r = (self.rank + k) % self.world
                xh, vh = everyone[r][slot]
                mapped.append((eng.peer_open(xh), eng.peer_open(vh)))
            self.peers.append(mapped)
        dist.barrier(group=group)

    def views(self, slot: int):
        """(xyz_all, valid_all) of this rank's own assembled buffers"""
        return self._views[slot]

    def select(self, slot: int):
        """Make slot the destination of the engine's next slr_run_mf / slr_run_ge calls."""
        targets = [self.own[slot]] + self.peers[slot]
        self.eng.set_gather_targets([t[0] for t in targets], [t[1] for t in targets], self.rank * self.b)

    def close(self):
        self.eng.set_gather_targets([], [])
        for mapped in self.peers:
            for xp, vp in mapped:
                self.eng.peer_close(xp)
                self.eng.peer_close(vp)
        self._views = []
        for xp, vp in self.own:
            self.eng.peer_free(xp)
            self.eng.peer_free(vp)
        self.peers, self.own = [], []


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


class CloudAssembly:
    """The assembled cloud of all ranks, allocated ONCE: [world * b_local, H, W, 3] xyz + [world * b_local, H, W] valid.

    Every rank's kernels write straight into its own block (`local_views()` are views of the assembled tensors), and
    `gather()` is an in-place all-gather (send buffer = this rank's block of the receive buffer), so assembling the
    cloud costs no copy besides the NVLink transfer itself.  With `slots=2` the gather of step k runs on a side
    stream while step k+1 computes into the other slot (events order producer and consumer both ways).
    Equal shards only (b_local scans on every rank); ragged batches go through `gather_clouds`."""

    def __init__(self, b_local: int, H: int, W: int, device, slots: int = 2, group=None, native=None):
        # native = (engine, ncclComm_t from engine.nccl_comm_create): gather through the C ABI (slr_allgather, NCCL bound
        # by libslr_b200.so itself) instead of torch.distributed's collective
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.b = b_local
        self.native = native
        n = self.world * b_local
        self.xyz = [torch.empty((n, H, W, 3), dtype=torch.float32, device=device) for _ in range(slots)]
        self.valid = [torch.empty((n, H, W), dtype=torch.uint8, device=device) for _ in range(slots)]
        self.cuda = torch.device(device).type == "cuda"
        if self.cuda:
            self.comm = torch.cuda.Stream(device=device)
            self.filled = [torch.cuda.Event() for _ in range(slots)]    # compute -> comm
            self.gathered = [torch.cuda.Event() for _ in range(slots)]  # comm -> compute (slot may be overwritten)
            self.used = [False] * slots

    def local_views(self, slot: int):
        lo = self.rank * self.b
        return self.xyz[slot][lo:lo + self.b], self.valid[slot][lo:lo + self.b]

    def before_compute(self, slot: int):
        """Call before launching kernels that write slot: waits until the previous gather of that slot is done."""
        if self.cuda and self.used[slot]:
            self.torch.cuda.current_stream().wait_event(self.gathered[slot])

    def gather(self, slot: int):
        """In-place all-gather of slot (asynchronous on GPUs: ordered after the kernels already launched on the
        current stream, executed on the communication stream)."""
        xl, vl = self.local_views(slot)
        if not self.cuda:
            self.dist.all_gather_into_tensor(self.xyz[slot], xl, group=self.group)
            self.dist.all_gather_into_tensor(self.valid[slot], vl, group=self.group)
            return self.xyz[slot], self.valid[slot]
        torch = self.torch
        self.filled[slot].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.filled[slot])
            if self.native is not None:
                eng, comm = self.native   # the engine binds to the current (= communication) stream for this call
                eng.allgather(comm, self.world, self.rank, self.b, self.xyz[slot], self.valid[slot])
            else:
                self.dist.all_gather_into_tensor(self.xyz[slot], xl, group=self.group)
                self.dist.all_gather_into_tensor(self.valid[slot], vl, group=self.group)
            self.gathered[slot].record(self.comm)
        self.used[slot] = True
        return self.xyz[slot], self.valid[slot]

    def wait(self):
        """Make the current stream wait for every gather issued so far."""
        if self.cuda:
            self.torch.cuda.current_stream().wait_stream(self.comm)
