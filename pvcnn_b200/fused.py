"""Host side of the fused PVConv block: one C-ABI call forward (pvcnn_pvconv_forward) and one backward
(pvcnn_pvconv_backward).  torch is used for device memory, streams and autograd plumbing only.

Reference being replaced: modules/pvconv.py:33-39 (forward wiring) + torch autograd of
modules/voxelization.py, modules/functional/{voxelization,devoxelization}.py, nn.Conv3d / nn.BatchNorm3d /
nn.LeakyReLU / nn.Conv1d / nn.BatchNorm1d / nn.ReLU.
"""
import ctypes
import os

import torch
from torch.autograd import Function

from . import _lib

_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


class Desc(ctypes.Structure):
    _fields_ = [("b", ctypes.c_int), ("n", ctypes.c_int), ("cin", ctypes.c_int), ("cout", ctypes.c_int),
                ("r", ctypes.c_int), ("normalize", ctypes.c_int), ("eps", ctypes.c_float),
                ("training", ctypes.c_int), ("npass", ctypes.c_int), ("bn_eps_vox", ctypes.c_float),
                ("bn_eps_pt", ctypes.c_float), ("momentum", ctypes.c_float), ("slope", ctypes.c_float),
                ("with_se", ctypes.c_int), ("vox_stats", ctypes.c_int), ("prepared", ctypes.c_int)]


_PARAM_FIELDS = ["w1", "b1", "g1", "be1", "rm1", "rv1", "w2", "b2", "g2", "be2", "rm2", "rv2",
                 "wp", "bp", "gp", "bep", "rmp", "rvp", "se_w1", "se_w2"]
_GRAD_FIELDS = ["w1", "b1", "g1", "be1", "w2", "b2", "g2", "be2", "wp", "bp", "gp", "bep", "se_w1", "se_w2"]
_WS_FIELDS = [("nc", _F), ("vc", _I), ("ind", _I), ("cnt", _I), ("fcl", _F), ("fcl_lo", _F), ("g0", _F),
              ("g0_lo", _F), ("y1", _F), ("z1", _F), ("z1_lo", _F), ("y2", _F), ("p", _F), ("coef", _F),
              ("wprep", _F), ("partials", _F), ("sums", _F), ("ga", _F), ("gpp", _F), ("gpp_lo", _F),
              ("gfpt", _F), ("d2", _F), ("gy2", _F), ("gy2_lo", _F), ("gy1", _F), ("gy1_lo", _F), ("sparse", _I), ("se", _F),
              ("vox_mean", _F), ("vox_denom", _F), ("prep", _F)]


class Params(ctypes.Structure):
    _fields_ = [(k, _F) for k in _PARAM_FIELDS] + [(k, ctypes.c_void_p) for k in ("nbt1", "nbt2", "nbtp")]


class Grads(ctypes.Structure):
    _fields_ = [(k, _F) for k in _GRAD_FIELDS]


class Workspace(ctypes.Structure):
    _fields_ = _WS_FIELDS


def _ptr(t, typ=_F):
    return ctypes.cast(ctypes.c_void_p(0 if t is None else t.data_ptr()), typ)


def precision_passes():
    """PVCNN_B200_PRECISION=fp32 (default, 3xTF32, parity mode) | tf32 (single pass, like the
    reference's cuDNN default)."""
    return 1 if os.environ.get("PVCNN_B200_PRECISION", "fp32").lower() == "tf32" else 3


def _pad4(x):
    return (x + 3) // 4 * 4


_SCRATCH = {}


def _scratch(name, numel, device, dtype=torch.float32):
    """Backward-only scratch is shared by every PVConv block on the device (grown on demand)."""
    key = (name, str(device), dtype)
    t = _SCRATCH.get(key)
    if t is None or t.numel() < numel:
        t = torch.empty(int(numel), dtype=dtype, device=device)
        _SCRATCH[key] = t
    return t


def _poison(tensors):
    """PVCNN_B200_POISON=1 (debug): fill every workspace buffer with NaN / INT_MAX before use, so that a kernel
    reading memory it never wrote shows up as NaN instead of depending on what the allocator handed out."""
    if os.environ.get("PVCNN_B200_POISON", "0") != "1":
        return
    for t in tensors:
        if t is not None:
            t.fill_(float("nan") if t.is_floating_point() else 2 ** 31 - 1)


class _Plan:
    """Buffers of one forward pass (kept alive for the backward)."""

    def __init__(self, desc, device, need_backward):
        lib = _lib.load()
        lib.pvcnn_pvconv_wprep_floats.restype = ctypes.c_longlong
        lib.pvcnn_pvconv_partials_floats.restype = ctypes.c_longlong
        lib.pvcnn_pvconv_sparse_ints.restype = ctypes.c_longlong
        b, n, r = desc.b, desc.n, desc.r
        ci, co = _pad4(desc.cin), _pad4(desc.cout)
        mv, mp = b * r ** 3, b * n
        lo = desc.npass > 1
        self.grid_lo = bool(lib.pvcnn_pvconv_needs_grid_lo(ctypes.byref(desc)))
        f = lambda numel: torch.empty(int(numel), dtype=torch.float32, device=device)
        i = lambda numel: torch.empty(int(numel), dtype=torch.int32, device=device)
        alloc = {}
        saved = dict(nc=b * 3 * n, fcl=mp * ci, g0=mv * ci, y1=mv * co, z1=mv * co, y2=mv * co, p=mp * co,
                     coef=12 * co)
        if lo:
            saved.update(fcl_lo=mp * ci)
        if self.grid_lo:
            saved.update(g0_lo=mv * ci, z1_lo=mv * co)
        if desc.with_se:
            saved.update(se=b * (7 * co + desc.cout // 8))
        for k, v in saved.items():
            # activations needed by the backward belong to this call; in inference they are scratch
            alloc[k] = f(v) if need_backward else _scratch("fwd_" + k, v, device)
        alloc["nc"] = f(b * 3 * n)
        alloc["vc"] = _scratch("vc", b * 3 * n, device, torch.int32)
        alloc["ind"] = i(b * n) if need_backward else _scratch("ind", b * n, device, torch.int32)
        alloc["cnt"] = i(b * r ** 3) if need_backward else _scratch("cnt", b * r ** 3, device, torch.int32)
        alloc["wprep"] = _scratch("wprep", lib.pvcnn_pvconv_wprep_floats(ctypes.byref(desc)), device)
        alloc["partials"] = _scratch("partials", lib.pvcnn_pvconv_partials_floats(ctypes.byref(desc)), device)
        alloc["sums"] = _scratch("sums", 16 * max(ci, co), device)
        nsp = lib.pvcnn_pvconv_sparse_ints(ctypes.byref(desc))  # activity lists: built in forward, reused in backward
        alloc["sparse"] = i(nsp) if need_backward else _scratch("sparse", nsp, device, torch.int32)
        self.t = alloc
        self.desc = desc
        self.device = device
        self.private = bool(need_backward)
        _poison(alloc.values())

    def add_backward_scratch(self):
        d = self.desc
        ci, co = _pad4(d.cin), _pad4(d.cout)
        mv, mp = d.b * d.r ** 3, d.b * d.n
        lo = d.npass > 1
        sizes = dict(ga=mp * co, gpp=mp * co, gfpt=mp * ci, d2=mv * max(ci, co), gy2=mv * co, gy1=mv * co)
        if lo:
            sizes.update(gpp_lo=mp * co)
        if self.grid_lo:
            sizes.update(gy2_lo=mv * co, gy1_lo=mv * co)
        for k, v in sizes.items():
            self.t[k] = _scratch("bwd_" + k, v, self.device)
        _poison(self.t[k] for k in sizes)

    def struct(self):
        ws = Workspace()
        for name, typ in _WS_FIELDS:
            setattr(ws, name, _ptr(self.t.get(name), typ))
        return ws


def _module_tensors(m):
    c1, n1, c2, n2 = m.voxel_layers[0], m.voxel_layers[1], m.voxel_layers[3], m.voxel_layers[4]
    cp, npt = m.point_features.layers[0], m.point_features.layers[1]
    return [c1.weight, c1.bias, n1.weight, n1.bias, c2.weight, c2.bias, n2.weight, n2.bias,
            cp.weight, cp.bias, npt.weight, npt.bias], [n1, n2, npt]


class _PVConvFused(Function):
    @staticmethod
    def forward(ctx, features, coords, module, need_bwd, w1, b1, g1, be1, w2, b2, g2, be2, wp, bp, gp, bep, se_w1=None,
                se_w2=None):
        dev = features.device
        if dev.type != "cuda":
            raise RuntimeError("features must be a CUDA tensor")  # utils.hpp:7 semantics: no CPU path
        features = features.contiguous().float()
        coords = coords.detach().contiguous().float()
        b, cin, n = features.shape
        training = bool(module.training)
        vox = module.voxelization
        desc = Desc(b, n, cin, module.out_channels, int(module.resolution), int(bool(vox.normalize)), float(vox.eps),
                    int(training), precision_passes(), 1e-4, 1e-5, 0.1, 0.1, int(se_w1 is not None), 0, 0)
        bns = [module.voxel_layers[1], module.voxel_layers[4], module.point_features.layers[1]]
        desc.bn_eps_vox = float(bns[0].eps)
        desc.bn_eps_pt = float(bns[2].eps)
        desc.momentum = float(bns[0].momentum if bns[0].momentum is not None else 0.1)
        desc.slope = float(module.voxel_layers[2].negative_slope)
        # need_bwd is decided by the caller: grad mode is always off inside Function.forward
        # per-cloud coordinate mean from the reference's own ATen reduction (bit-exact voxel indices)
        from .functional.ops import voxel_stats
        desc.vox_stats, vmean, vdenom = voxel_stats(coords, bool(vox.normalize), float(vox.eps))
        plan = _Plan(desc, dev, bool(need_bwd) and training)
        plan.t["vox_mean"], plan.t["vox_denom"] = vmean, vdenom
        prm = Params()
        vals = dict(w1=w1, b1=b1, g1=g1, be1=be1, rm1=bns[0].running_mean, rv1=bns[0].running_var,
                    w2=w2, b2=b2, g2=g2, be2=be2, rm2=bns[1].running_mean, rv2=bns[1].running_var,
                    wp=wp, bp=bp, gp=gp, bep=bep, rmp=bns[2].running_mean, rvp=bns[2].running_var, se_w1=se_w1, se_w2=se_w2)
        keep = []
        for k in _PARAM_FIELDS:
            t = vals[k]
            if t is not None:
                t = t.detach()
                if not t.is_contiguous() or t.dtype != torch.float32:
                    raise RuntimeError("PVConv parameters must be contiguous float32 tensors")
            keep.append(t)
            setattr(prm, k, _ptr(t))
        # nn.BatchNorm's step counters are incremented by the statistics kernels (no ATen launch per layer and step)
        for name, bn in zip(("nbt1", "nbt2", "nbtp"), bns):
            t = bn.num_batches_tracked
            ok = training and t is not None and t.is_cuda and t.dtype == torch.int64
            setattr(prm, name, t.data_ptr() if ok else None)
            if training and t is not None and not ok:
                t += 1
        out = torch.empty((b, module.out_channels, n), dtype=torch.float32, device=dev)
        _poison([out])
        if not training and not torch.is_grad_enabled() and os.environ.get("PVCNN_B200_PVCONV_EVAL", "cached") != "rebuild":
            # frozen block: GEMM operands, BatchNorm coefficients and the conv2 constants are rebuilt only when a parameter
            # or running statistic changed (in-place updates bump the version counter, re-assignment changes the pointer)
            tensors = [t for t in keep[:18] if t is not None]
            key = tuple((t.data_ptr(), t._version) for t in tensors) + (str(dev), float(desc.bn_eps_vox),
                                                                       float(desc.bn_eps_pt), float(desc.slope))
            cache = getattr(module, "_pvcnn_eval_prep", None)
            if cache is None or cache[0] != key:
                lib = _lib.load()
                lib.pvcnn_pvconv_prep_floats.restype = ctypes.c_longlong
                prep = torch.empty(lib.pvcnn_pvconv_prep_floats(ctypes.byref(desc)), dtype=torch.float32, device=dev)
                module._pvcnn_eval_prep = (key, prep)
                desc.prepared = 0
            else:
                prep = cache[1]
                desc.prepared = 1
            plan.t["prep"] = prep
        ws = plan.struct()
        _lib.call("pvcnn_pvconv_forward", ctypes.byref(desc), features, coords, ctypes.byref(prm), ctypes.byref(ws),
                  out, device=dev)
        ctx.plan, ctx.prm, ctx.keep, ctx.desc, ctx.module = plan, prm, keep, desc, module
        ctx.shapes = [None if t is None else t.shape
                      for t in (w1, b1, g1, be1, w2, b2, g2, be2, wp, bp, gp, bep, se_w1, se_w2)]
        ctx.in_shape = features.shape
        return out

    @staticmethod
    def backward(ctx, grad_out):
        plan, desc = ctx.plan, ctx.desc
        if not desc.training:
            raise RuntimeError("PVConv backward requires training mode (batch statistics)")
        if not plan.private:
            raise RuntimeError("PVConv forward ran without saving activations (no_grad / eval); cannot run backward")
        dev = grad_out.device
        grad_out = grad_out.contiguous().float()
        plan.add_backward_scratch()
        nparam = 14 if ctx.shapes[12] is not None else 12  # SE weights are optional inputs of forward()
        # data-parallel bucket attached (pvcnn_b200/parallel.py): the kernels write the parameter gradients straight into
        # the flat all-reduce buffer (p.grad is a view of it) and autograd gets None for them
        views = getattr(ctx.module, "_pvcnn_grad_views", None)
        direct = views is not None and len(views) == nparam and views[0].device == dev
        if direct:
            grads = list(views) + [None] * (14 - nparam)
        else:
            grads = [None if s is None else torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.shapes]
        gs = Grads()
        for k, t in zip(_GRAD_FIELDS, grads):
            setattr(gs, k, _ptr(t))
        gfeat = torch.empty(ctx.in_shape, dtype=torch.float32, device=dev)
        _poison([gfeat] + ([] if direct else grads))
        ws = plan.struct()
        if direct:
            # phase 1 ends with the last parameter gradient; the bucket may launch its all-reduce on a side stream now,
            # overlapping phase 2 (conv1 data gradient + scatter back to the points)
            _lib.call("pvcnn_pvconv_backward_phase", ctypes.byref(desc), grad_out, ctypes.byref(ctx.prm),
                      ctypes.byref(ws), gfeat, ctypes.byref(gs), 1, device=dev)
            ctx.module._pvcnn_bucket.module_grads_ready(ctx.module)
            _lib.call("pvcnn_pvconv_backward_phase", ctypes.byref(desc), grad_out, ctypes.byref(ctx.prm),
                      ctypes.byref(ws), gfeat, ctypes.byref(gs), 2, device=dev)
            return (gfeat, None, None, None, *([None] * nparam))
        _lib.call("pvcnn_pvconv_backward", ctypes.byref(desc), grad_out, ctypes.byref(ctx.prm), ctypes.byref(ws),
                  gfeat, ctypes.byref(gs), device=dev)
        return (gfeat, None, None, None, *grads[:nparam])


def pvconv_fused(module, features, coords):
    """features [B,Cin,N], coords [B,3,N] -> fused features [B,Cout,N] (module: pvcnn_b200.nn.PVConv)."""
    params, _ = _module_tensors(module)
    if module.with_se:
        se = module.voxel_layers[6]
        params = params + [se.fc[0].weight, se.fc[2].weight]
    # activations saved for the backward must be private to this call; otherwise they live in shared scratch
    need_bwd = module.training and torch.is_grad_enabled() and (
        features.requires_grad or any(p.requires_grad for p in params))
    if not torch.is_grad_enabled():   # inference: plain call, no autograd.Function bookkeeping
        from .mlp import _NoCtx
        return _PVConvFused.forward(_NoCtx(), features, coords, module, need_bwd, *params)
    return _PVConvFused.apply(features, coords, module, need_bwd, *params)
