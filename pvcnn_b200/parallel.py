"""Data-parallel plumbing for the PVConv path: one process per GPU, ONE collective per step.

The reference trains with single-process `nn.DataParallel` (train.py:180-181): replicas see a batch
shard, BatchNorm statistics stay per replica, gradients are summed onto GPU 0.  Here every rank owns
one B200, keeps a replica, and all-reduces a single flat fp32 gradient bucket over NCCL (NVLink 5 /
NVSwitch).  BatchNorm buffers are not synchronised (same semantics as the reference's replicas).
"""
import torch
import torch.distributed as dist


class GradBucket:
    """Flat fp32 buffer aliasing every parameter gradient -> a single all-reduce per step."""

    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        device = device if device is not None else self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def pack(self):
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)

    def unpack(self):
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)

    def all_reduce_mean(self, group=None):
        """sum over ranks / world_size, written back into p.grad.  No-op without a process group."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self.pack()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(dist.get_world_size(group))
        self.unpack()


def shard_batch(tensor, rank, world):
    """Contiguous batch shard of rank `rank` (samples are independent units)."""
    b = tensor.shape[0]
    per = (b + world - 1) // world
    return tensor[rank * per:min(b, (rank + 1) * per)]


def broadcast_parameters(module, src=0):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)
