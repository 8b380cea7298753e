"""CPU: the multi-GPU host logic over gloo, world_size 2 and 3 (no GPU, no kernels: a stand-in local tracker)."""
import os
import socket

import numpy as np
import pytest

from mtf_b200 import sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 1000, 1024, 8192):
        for w in (1, 2, 3, 4, 8):
            got = [sharding.shard_range(n, w, r) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == sharding.shard_sizes(n, w)


class _FakeTracker:
    """moves every corner by +1 px per update: enough to see ordering mistakes in the gather"""

    def __init__(self, n):
        self.n = n

    def initialize(self, corners, img):
        assert corners.shape == (self.n, 2, 4)
        self.c = corners.copy()

    def update(self, img):
        self.c += 1.0

    def getRegion(self):
        return self.c


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        corners = np.arange(n_total * 8, dtype=np.float64).reshape(n_total, 2, 4)
        tr = sharding.ShardedBatchTracker(n_total, lambda n: _FakeTracker(n))
        tr.initialize(corners, None)
        tr.update(None); tr.update(None)
        out = tr.getRegion().numpy()
        q.put((rank, bool(np.array_equal(out, corners + 2.0)), tr.lo, tr.hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 64), (2, 37), (3, 10)])
def test_sharded_tracker_gloo(world, n_total):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    spans = sorted((lo, hi) for _, _, lo, hi in res)
    assert spans[0][0] == 0 and spans[-1][1] == n_total


def _handle_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h = np.full(64, rank + 1, dtype=np.uint8); h[0] = 200 + rank
        out = sharding.exchange_peer_handles(h)
        q.put((rank, out.shape, out[:, 0].tolist(), out[:, 1].tolist()))
    finally:
        dist.destroy_process_group()


def test_peer_handles_travel_in_rank_order():
    """the 64-byte IPC handles of mtfb_peer_export are all-gathered by whatever backend the group has (here gloo, two and three
    ranks): row r of the result is rank r's handle on every rank -- what mtfb_peer_attach expects"""
    import torch.multiprocessing as mp
    for world in (2, 3):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_handle_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for rank, shape, first, second in res:
            assert shape == (world, 64)
            assert first == [200 + r for r in range(world)] and second == [r + 1 for r in range(world)]
