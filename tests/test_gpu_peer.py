"""GPU: the per-frame exchange of the corners over peer memory (mtfb_peer_*, peer_gather.cu).

One GPU: a job of one rank (the kernel's fused stores into the gathered array, the signal / wait kernel, the double
buffering, the push of corners that no update kernel wrote, the lock-step rule).  Two or more GPUs: two processes, the IPC
handles exchanged over gloo, every rank's gathered array against the other rank's own corners -- skipped on a one-GPU box
(bench.py --gpus N checks the same exchange against NCCL's all-gather on every multi-GPU run)."""
import os
import socket

import numpy as np
import pytest

from common import patches

pytestmark = pytest.mark.gpu


def _tracker(n, **kw):
    from mtf_b200 import api
    return api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=n, **kw))


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_single_rank_job(seq384, precision):
    from mtf_b200 import api
    frames, _ = seq384
    cs = patches(12, 49.0, 384, 384)
    n_total, row0 = 20, 5                                   # the rank's rows sit in the middle of a larger job
    tr = _tracker(len(cs), precision=precision)
    ref = _tracker(len(cs), precision=precision)
    h = tr.peer_export(n_total)
    assert h.shape == (64,) and h.any()
    tr.peer_attach(0, 1, row0, h.reshape(1, 64))
    tr.initialize(cs, frames[0]); ref.initialize(cs, frames[0])
    tr.peer_gather()                                        # corners of initialize(): pushed by the small kernel
    g = tr.getGatheredRegion()
    assert np.array_equal(g[row0:row0 + len(cs)], tr.getRegion()) and not g[:row0].any() and not g[row0 + len(cs):].any()
    for k in (1, 2, 3):
        tr.update(frames[k]); ref.update(frames[k])
        tr.peer_gather()
        g = tr.getGatheredRegion()
        assert np.array_equal(g[row0:row0 + len(cs)], ref.getRegion())      # the update kernel's own stores, both buffers in turn
        assert np.array_equal(tr.getRegion(), ref.getRegion())
    tr.update(frames[1])
    with pytest.raises(api.MTFError):                       # the ranks of a job exchange every frame
        tr.update(frames[2])
    tr.peer_gather()
    tr.setRegion(cs)
    tr.peer_gather()
    assert np.array_equal(tr.getGatheredRegion()[row0:row0 + len(cs)], cs)


def test_peer_argument_checks(seq384):
    from mtf_b200 import api
    tr = _tracker(4)
    with pytest.raises(api.MTFError):
        tr.peer_gather()                                    # not attached
    h = tr.peer_export(4)
    with pytest.raises(api.MTFError):
        tr.peer_export(4)                                   # twice
    with pytest.raises(api.MTFError):
        tr.peer_attach(0, 9, 0, np.zeros((9, 64), np.uint8))
    with pytest.raises(api.MTFError):
        tr.peer_attach(0, 1, 2, h.reshape(1, 64))           # rows 2 .. 6 of 4


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from mtf_b200 import api, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(rank)
        frames, _ = synth.make_sequence(4, 384, 384, seed=1234, walk_seed=5678, sigma=1.0)
        cs = patches(23, 49.0, 384, 384)
        sh = sharding.ShardedBatchTracker(len(cs), lambda n: api.BatchTracker(api.make_params(
            "ssd", "homography", "fclk", n_patches=n, precision="f32", device=rank)))
        sh.attach_peers()
        sh.initialize(cs, frames[0])
        whole = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=len(cs), precision="f32", device=rank))
        whole.initialize(cs, frames[0])
        ok = True
        for k in (1, 2, 3):
            sh.update(frames[k]); whole.update(frames[k])
            ok = ok and bool(np.array_equal(sh.getRegion(), whole.getRegion()))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_ipc():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
