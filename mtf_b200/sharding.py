"""Multi-GPU host logic: independent patches / grid cells / particles shard embarrassingly.

One process per GPU.  Every rank owns a contiguous index range of the batch, tracks it with its own
BatchTracker on its own copy of the frame, and the per-patch results (P_local x 8 corners, or particle
weights) are all-gathered once per frame -- LK iterations of different patches never interact, so no
collective is needed inside the iteration loop.  The reference has no counterpart: its fan-out is a
shared-memory loop over `trackers[i]->update()` (SM/src/GridTracker.cc:247-264)."""
import numpy as np


def shard_range(n_items, world_size, rank):
    """contiguous range [lo, hi) of rank: sizes differ by at most one, earlier ranks take the remainder"""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items, world_size):
    return [shard_range(n_items, world_size, r)[1] - shard_range(n_items, world_size, r)[0] for r in range(world_size)]


def all_gather_rows(local, n_total, group=None, out=None):
    """all-gather a (n_local, ...) tensor whose row counts follow shard_range into a (n_total, ...) tensor.

    Uses all_gather_into_tensor when the shards are equal (one NCCL call on the caller's stream; `out`, if given, receives
    the result: a per-frame caller keeps one buffer), padded all_gather otherwise."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    tail = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        if out is None:
            out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
        return out
    m = max(sizes)
    padded = torch.zeros((m,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def device_view(ptr, shape, device, typestr="<f8"):
    """zero-copy torch view of a device array the library owns (mtfb_device_results): nothing is copied or moved"""
    import torch

    class _Raw:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Raw(), device=device)


def exchange_peer_handles(handle, group=None):
    """all-gather the ranks' 64-byte IPC handles (host side, through whatever backend the group has)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    t = torch.from_numpy(np.ascontiguousarray(handle, dtype=np.uint8))
    if backend == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return np.stack([p.cpu().numpy() for p in parts])


class ShardedBatchTracker:
    """The batch of `corners.shape[0]` patches split over the ranks of a torch.distributed group.

    make_local(n_local) must return an object with initialize / update / getRegion (a BatchTracker); it is a
    parameter so that the sharding logic can be exercised without a GPU."""

    def __init__(self, n_total, make_local, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_total = n_total
        self.lo, self.hi = shard_range(n_total, self.world, self.rank)
        self.local = make_local(self.hi - self.lo)
        self._d_corners = None
        self._gathered = None
        self._peer = False

    def attach_peers(self):
        """set up the exchange over NVLink peer memory (mtfb_peer_*): from now on update() writes every patch's corners into
        the gathered arrays of all ranks and getRegion(device=...) needs no collective call"""
        h = self.local.peer_export(self.n_total)
        handles = exchange_peer_handles(h, self.group)
        self.local.peer_attach(self.rank, self.world, self.lo, handles)
        self._peer = True

    def gathered_device(self, device):
        """(n_total, 8) torch view of the current gathered array after the frame's signal / wait"""
        self.local.peer_gather()
        ptr, n = self.local.peer_gathered_ptr()
        return device_view(ptr, (n, 8), device)

    def initialize(self, corners, img):
        c = np.asarray(corners, dtype=np.float64).reshape(self.n_total, 2, 4)
        self.local.initialize(c[self.lo:self.hi], img)

    def update(self, img):
        self.local.update(img)

    def local_corners_device(self, device):
        """the local tracker's (n_local, 8) result array where the kernel wrote it: a torch view of the library's device
        buffer (mtfb_device_results), no host hop"""
        if self._d_corners is None:
            ptr = self.local.device_results()[0]
            self._d_corners = device_view(ptr, (self.hi - self.lo, 8), device)
        return self._d_corners

    def getRegion(self, device=None):
        """(n_total, 2, 4) corners of every patch, on every rank.  With a CUDA `device` and a local tracker that exposes its
        device result array (BatchTracker.device_results) the all-gather reads the kernel's output buffer directly (NCCL on
        the current stream, which must be the tracker's stream: bench.py sets both); otherwise (CPU tests over gloo) through
        the host getter."""
        import torch
        if self._peer:
            if device is not None:
                return self.gathered_device(device).reshape(self.n_total, 2, 4)
            self.local.peer_gather()
            return self.local.getGatheredRegion()
        if device is not None and hasattr(self.local, "device_results"):
            loc = self.local_corners_device(device)
            if self._gathered is None:
                self._gathered = torch.empty((self.n_total, 8), dtype=torch.float64, device=device)
            return all_gather_rows(loc, self.n_total, self.group, out=self._gathered).reshape(self.n_total, 2, 4)
        else:
            loc = torch.as_tensor(np.ascontiguousarray(self.local.getRegion()).reshape(-1, 8))
            if device is not None:
                loc = loc.to(device)
        return all_gather_rows(loc, self.n_total, self.group).reshape(self.n_total, 2, 4)
