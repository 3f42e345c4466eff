"""world_size-2 gloo test of the N>1 host logic (flat gradient bucket, batch sharding)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pvcnn_b200.parallel import GradBucket, shard_batch, broadcast_parameters
    torch.manual_seed(100 + rank)                       # different init per rank ...
    m = torch.nn.Sequential(torch.nn.Conv1d(4, 8, 1), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Conv1d(8, 3, 1))
    broadcast_parameters(m)                             # ... made identical by the broadcast
    torch.manual_seed(7)
    x = torch.randn(6, 4, 16)
    y = torch.randn(6, 3, 16)
    xs, ys = shard_batch(x, rank, world), shard_batch(y, rank, world)
    bucket = GradBucket(m.parameters())
    loss = ((m(xs) - ys) ** 2).sum()
    loss.backward()
    local = [p.grad.clone() for p in m.parameters()]
    bucket.all_reduce_mean()
    got = [p.grad.clone() for p in m.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, [g.numpy() for g in local])
    if rank == 0:
        import numpy as np
        for i, g in enumerate(got):
            want = sum(gathered[r][i] for r in range(world)) / world
            assert np.allclose(g.numpy(), want, rtol=1e-6, atol=1e-7)
        w0 = [p.detach().clone() for p in m.parameters()]
    # round-2 path: p.grad ARE views of the flat buffer (autograd accumulates into them), zero() + finish() per step
    b2 = GradBucket(m.parameters()).attach(m)
    for _ in range(2):
        b2.zero()
        ((m(xs) - ys) ** 2).sum().backward()
        b2.finish()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(b2.params, b2.views))
    if rank == 0:
        import numpy as np
        for i, p in enumerate(m.parameters()):
            want = sum(gathered[r][i] for r in range(world)) / world
            assert np.allclose(p.grad.numpy(), want, rtol=1e-5, atol=1e-6)
        out.put(("ok", len(bucket.flat), float(sum(w.abs().sum() for w in w0))))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    status, nflat, _ = q.get(timeout=5)
    assert status == "ok" and nflat == 4 * 8 + 8 + 8 + 8 + 8 * 3 + 3


def test_shard_batch_covers_everything():
    from pvcnn_b200.parallel import shard_batch
    x = torch.arange(16).view(16, 1)
    for world in (1, 2, 4, 8):
        parts = [shard_batch(x, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts), x)
    assert shard_batch(torch.arange(5).view(5, 1), 2, 4).numel() == 1  # ragged tail


def _eval_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pvcnn_b200 import evaluate as E
    num_scenes, classes = 7, 13
    mine = E.scenes_of_rank(num_scenes)
    stats = torch.zeros(3, classes, num_scenes, dtype=torch.int64)
    for s in mine:   # stand-in for SceneVotes.stats of scene s (the kernels need a GPU; the exchange step does not)
        g = torch.Generator().manual_seed(1000 + s)
        stats[:, :, s] = torch.randint(0, 10_000, (3, classes), generator=g)
    E.all_reduce_stats(stats)
    if rank == 0:
        out.put((mine, stats))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluation_scenes_and_stats_exchange_gloo_world2():
    """evaluation shards over whole scenes (round-robin), one SUM all-reduce of the [3, classes, scenes] counters"""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    mine, stats = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mine == [0, 2, 4, 6]
    for s in range(7):
        g = torch.Generator().manual_seed(1000 + s)
        assert torch.equal(stats[:, :, s], torch.randint(0, 10_000, (3, 13), generator=g))
    from pvcnn_b200 import evaluate as E
    assert E.scenes_of_rank(5, 1, 3) == [1, 4] and E.scenes_of_rank(2, 0, 1) == [0, 1]
