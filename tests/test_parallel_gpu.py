"""Hardware multi-GPU correctness (SURVEY.md 4 "Distributed"): two ranks, one B200 each, NCCL.  The flat gradient bucket
that the fused PVConv backward writes into, all-reduced (AVG) on the side stream between the two backward phases, must
equal the mean of the per-shard single-GPU gradients.  Skipped on a 1-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import modules
    from pvcnn_b200.parallel import GradBucket, shard_batch, broadcast_parameters
    torch.manual_seed(100 + rank)
    m = modules.PVConv(16, 32, 3, 16).to(dev).train()
    broadcast_parameters(m)
    g = torch.Generator().manual_seed(5)
    b, n = 4, 2048
    f = torch.randn(b, 16, n, generator=g)
    c = torch.rand(b, 3, n, generator=g)
    go = torch.randn(b, 32, n, generator=g)
    fs, cs, gs = [shard_batch(t, rank, world).contiguous().to(dev) for t in (f, c, go)]
    # (1) plain single-GPU backward on the shard: local gradients through autograd
    fl = fs.clone().requires_grad_(True)
    out, _ = m((fl, cs))
    out.backward(gs)
    local = [p.grad.detach().clone() for p in m.parameters()]
    for p in m.parameters():
        p.grad = None
    # (2) bucket path: gradients written into the flat buffer by the kernels, all-reduce launched inside the backward
    bucket = GradBucket(list(m.parameters()), dev).attach(m)
    assert len(bucket._direct_modules) == 1
    for _ in range(2):   # second step: the buffer is reused (overwrite semantics)
        bucket.zero()
        fb = fs.clone().requires_grad_(True)
        out_b, _ = m((fb, cs))
        out_b.backward(gs)
        bucket.finish()
    torch.cuda.synchronize()
    got = [p.grad.detach().clone() for p in m.parameters()]
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    gathered = [torch.zeros_like(torch.cat([t.flatten() for t in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([t.flatten() for t in local]))
    want = sum(gathered) / world
    have = torch.cat([t.flatten() for t in got])
    err = float((have - want).abs().max() / want.abs().max())
    err_in = float((fb.grad - fl.grad).abs().max() / fl.grad.abs().max())
    if rank == 0:
        q.put((err, err_in))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_bucket_allreduce_equals_mean_of_shard_grads_nccl_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    err, err_in = q.get(timeout=10)
    # atomics in the scatter / split-K kernels reorder fp32 sums between the two runs: rounding-level agreement
    assert err < 2e-5, err
    assert err_in < 2e-5, err_in
