"""Data-parallel plumbing for the PVConv path: one process per GPU, ONE collective per step.

The reference trains with single-process `nn.DataParallel` (train.py:180-181): replicas see a batch
shard, BatchNorm statistics stay per replica, gradients are summed onto GPU 0.  Here every rank owns
one B200, keeps a replica, and all-reduces a single flat fp32 gradient bucket over NCCL (NVLink 5 /
NVSwitch).  BatchNorm buffers are not synchronised (same semantics as the reference's replicas).

Round 2: the bucket is not a staging copy any more.  `p.grad` of every parameter IS a view of the flat buffer, the fused
PVConv backward writes its parameter gradients straight into those views (no autograd accumulation kernels, no pack /
unpack), the average is NCCL's own ReduceOp.AVG (no divide kernel), and the collective is launched on a side stream as
soon as the last parameter gradient of the step exists -- between the two phases of pvcnn_pvconv_backward_phase -- so
it overlaps the input-gradient kernels.
"""
import torch
import torch.distributed as dist


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class GradBucket:
    """Flat fp32 buffer that owns every parameter gradient -> a single all-reduce per step.

    attach(model): p.grad becomes a persistent view of the flat buffer; fused PVConv blocks are told to write their
    gradients there directly (overwrite semantics: one backward per step).  Parameters of other layers keep autograd's
    accumulate-into-.grad behaviour, so call zero() at the start of a step instead of setting .grad to None."""

    def __init__(self, params, device=None, group=None):
        self.params = [p for p in params if p.requires_grad]
        device = device if device is not None else self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=device)
        self.group = group
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._attached = False
        self._direct_modules = []
        self._pending = 0
        self._launched = False
        self._comm_stream = None
        self._done = None

    # ---- round-2 path: gradients live in the bucket -------------------------------------------------------------
    def attach(self, model):
        from .nn.pvconv import PVConv
        view_of = {id(p): v for p, v in zip(self.params, self.views)}
        for p in self.params:
            p.grad = view_of[id(p)]
        self._direct_modules = []
        for m in model.modules():
            if isinstance(m, PVConv) and m._fused_unsupported_reason() is None:
                from .fused import _module_tensors
                tensors, _ = _module_tensors(m)
                if m.with_se:
                    tensors = tensors + [m.voxel_layers[6].fc[0].weight, m.voxel_layers[6].fc[2].weight]
                if all(id(t) in view_of for t in tensors):
                    m._pvcnn_grad_views = [view_of[id(t)] for t in tensors]
                    m._pvcnn_bucket = self
                    self._direct_modules.append(m)
        self._attached = True
        if self.flat.is_cuda:
            self._comm_stream = torch.cuda.Stream(device=self.flat.device)
            self._done = torch.cuda.Event()
        return self

    def zero(self):
        """Start of a step (replaces `p.grad = None`): one memset; gradients of non-PVConv layers accumulate into it."""
        self._pending = len(self._direct_modules)
        self._launched = False
        only_direct = self._attached and len(self._direct_modules) > 0 and sum(
            v.numel() for m in self._direct_modules for v in m._pvcnn_grad_views) == self.flat.numel()
        if not only_direct:
            self.flat.zero_()   # fused PVConv blocks overwrite their slices; everything else accumulates

    def module_grads_ready(self, module):
        """Called by the fused PVConv backward between its two phases.  When the LAST direct-write block of the step
        reports and no other layer owns parameters, the collective starts right away on the side stream."""
        self._pending -= 1
        if self._pending == 0 and self._only_direct():
            self._launch()

    def _only_direct(self):
        return sum(v.numel() for m in self._direct_modules for v in m._pvcnn_grad_views) == self.flat.numel()

    def _launch(self):
        if self._launched or _world(self.group) == 1:
            self._launched = True
            return
        if not self.flat.is_cuda:                    # gloo (CPU tests): synchronous, no AVG op
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(_world(self.group))
            self._launched = True
            return
        cur = torch.cuda.current_stream(self.flat.device)
        self._comm_stream.wait_stream(cur)          # every parameter gradient has been enqueued on `cur`
        with torch.cuda.stream(self._comm_stream):
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)   # NCCL averages in the collective
            self._done.record(self._comm_stream)
        self._launched = True

    def finish(self):
        """End of the backward: make sure the collective has been launched and order the stream behind it."""
        if _world(self.group) == 1:
            return
        if not self._attached:
            return self.all_reduce_mean(self.group)
        if not self._launched:
            self._launch()
        if self.flat.is_cuda:
            torch.cuda.current_stream(self.flat.device).wait_event(self._done)

    # ---- round-1 path (staging copy), kept for CPU / gloo and for parameters that were never attached -------------
    def pack(self):
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)

    def unpack(self):
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.clone()
            elif p.grad.data_ptr() != v.data_ptr():
                p.grad.copy_(v)

    def all_reduce_mean(self, group=None):
        """sum over ranks / world_size, written back into p.grad.  No-op without a process group."""
        world = _world(group)
        if world == 1:
            return
        if self._attached:
            return self.finish()
        self.pack()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)
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


def pin_process_to_gpu_numa_node(device_index):
    """Bind this rank's CPU threads (and therefore its first-touch pinned host buffers) to the NUMA node its GPU hangs
    off, so that eight ranks' host<->device copies do not all cross the socket interconnect.  Best effort: returns the
    node or None."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:  # noqa: BLE001
        return None
    return None
