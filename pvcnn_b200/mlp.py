"""Host side of the native SharedMLP (reference: modules/shared_mlp.py:6-33): every (1x1 conv, BatchNorm, ReLU) layer
runs as one C-ABI call forward (pvcnn_mlp_layer_forward: tcgen05 GEMM + fused BN statistics / apply) and one backward
(pvcnn_mlp_layer_backward), on channels-last rows.  torch is used for device memory and autograd plumbing only.

Layouts: the module boundary keeps the reference's [B, C, N] / [B, C, M, U]; inside, activations are [rows, pad4(C)]
with rows = B*N (or B*M*U).  `pool_u` folds the set-abstraction max over the U neighbours (modules/pointnet.py:87) into
the last layer so the [B, C, M, U] activation is never written.
"""
import ctypes
import os

import torch
from torch.autograd import Function

from . import _lib
from .fused import _scratch, precision_passes

_LL = ctypes.c_longlong


def _pad4(x):
    return (x + 3) // 4 * 4


class _NoCtx:
    """Stand-in for the autograd context when nothing will be differentiated (no_grad / inference): the Function's
    forward runs as a plain call, without torch.autograd.Function.apply's bookkeeping (the small frustum networks are
    host-bound: ~180 launches in 3.5 ms)."""
    needs_input_grad = (False,) * 64

    def save_for_backward(self, *a):
        pass

    def mark_non_differentiable(self, *a):
        pass


def _run(fn, *args):
    if torch.is_grad_enabled():
        return fn.apply(*args)
    return fn.forward(_NoCtx(), *args)


def _lib_sizes():
    lib = _lib.load()
    lib.pvcnn_mlp_partials_floats.restype = ctypes.c_longlong
    lib.pvcnn_mlp_wprep_floats.restype = ctypes.c_longlong
    return lib


class _ToCL(Function):
    """[B, C, N] -> ([B*N, pad4(C)], lo) ; backward: channels-last gradient -> [B, C, N]."""

    @staticmethod
    def forward(ctx, x, want_lo):
        x = x.contiguous().float()
        b, c, n = x.shape
        cp = _pad4(c)
        xcl = torch.empty((b * n, cp), dtype=torch.float32, device=x.device)
        lo = torch.empty_like(xcl) if want_lo else None
        _lib.call("pvcnn_points_to_cl", b, c, n, x, xcl, lo)
        ctx.shape = (b, c, n)
        if lo is None:
            lo = xcl.new_empty(0)
        ctx.mark_non_differentiable(lo)
        return xcl, lo

    @staticmethod
    def backward(ctx, g, _glo):
        b, c, n = ctx.shape
        out = torch.empty((b, c, n), dtype=torch.float32, device=g.device)
        _lib.call("pvcnn_cl_to_points", b, c, n, g.contiguous(), out)
        return out, None


class _FromCL(Function):
    """[B*N, pad4(C)] -> [B, C, N]."""

    @staticmethod
    def forward(ctx, xcl, b, c, n):
        out = torch.empty((b, c, n), dtype=torch.float32, device=xcl.device)
        _lib.call("pvcnn_cl_to_points", b, c, n, xcl.contiguous(), out)
        ctx.shape = (b, c, n)
        return out

    @staticmethod
    def backward(ctx, g):
        b, c, n = ctx.shape
        gcl = torch.empty((b * n, _pad4(c)), dtype=torch.float32, device=g.device)
        _lib.call("pvcnn_points_to_cl", b, c, n, g.contiguous().float(), gcl, None)
        return gcl, None, None, None


def _select_cols(w, cols):
    """Columns `cols` = ((start, end), ...) of a k=1 conv weight [cout, cin, 1(,1)] as a contiguous [cout, sum] matrix."""
    w2 = w.reshape(w.shape[0], w.shape[1])
    return torch.cat([w2[:, a:b] for a, b in cols], dim=1).contiguous()


def _eval_prepared(conv, bn, cin, cout, dev, lib, cols=None):
    """(wprep, coef) of a frozen layer, rebuilt only when a parameter or running statistic changed (in-place updates bump
    torch's version counter; re-assignment changes the storage pointer).  `cols`: the input-channel ranges the GEMM keeps
    (the rest of the weight acts through a per-cloud bias)."""
    tensors = (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple((t.data_ptr(), t._version) for t in tensors) + (float(bn.eps), str(dev), cols)
    cache = getattr(bn, "_pvcnn_eval_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1], cache[2]
    wprep = torch.empty(lib.pvcnn_mlp_wprep_floats(cin, cout), dtype=torch.float32, device=dev)
    coef = torch.empty(4 * _pad4(cout), dtype=torch.float32, device=dev)
    weight = conv.weight.detach() if cols is None else _select_cols(conv.weight.detach(), cols)
    _lib.call("pvcnn_mlp_layer_prepare", cin, cout, float(bn.eps), weight, bn.weight.detach(),
              bn.bias.detach(), bn.running_mean, bn.running_var, wprep, coef, device=dev)
    bn._pvcnn_eval_cache = (key, wprep, coef)
    return wprep, coef


def _mlp_eval(x_cl, x_lo, meta, group_bias=None, group_rows=0, want_lo=False):
    """Frozen (eval-mode) stack: per layer one GEMM with bias + BatchNorm + ReLU (+ lo split) in its epilogue."""
    lib = _lib_sizes()
    dev = x_cl.device
    rows = x_cl.shape[0]
    npass, pool_u, widths = meta["npass"], meta["pool_u"], meta["widths"]
    convs, bns = meta["convs"], meta["bns"]
    cin = meta["cin"]
    x, xl = x_cl, (x_lo if npass > 1 else None)
    nl = len(widths)
    out = None
    for li, cout in enumerate(widths):
        conv, bn = convs[li], bns[li]
        co = _pad4(cout)
        last = li == nl - 1
        pool = pool_u if (last and pool_u) else 0
        wprep, coef = _eval_prepared(conv, bn, cin, cout, dev, lib, meta.get("point_cols") if li == 0 else None)
        y = z = zl = pooled = argmax = tmp = None
        if pool:
            groups = rows // pool
            y = _scratch("mlp_y0", rows * co, dev)
            pooled = torch.empty((groups, co), dtype=torch.float32, device=dev)
            argmax = torch.empty((groups, co), dtype=torch.int32, device=dev)
            segs = lib.pvcnn_mlp_pool_segments(_LL(groups), pool)
            if segs > 1:
                tmp = _scratch("mlp_pooltmp", 2 * groups * segs * co, dev)
        else:
            z = torch.empty((rows, co), dtype=torch.float32, device=dev)
            zl = torch.empty((rows, co), dtype=torch.float32, device=dev) if (npass > 1 and (not last or want_lo)) else None
        gb = group_bias if li == 0 else None
        _lib.call("pvcnn_mlp_layer_forward_eval", _LL(rows), cin, cout, npass, x, xl, wprep,
                  None if conv.bias is None else conv.bias.detach(), coef, _LL(group_rows if gb is not None else 0), gb,
                  0 if gb is None else gb.shape[1], y, z, zl, pool, pooled, argmax, tmp, device=dev)
        out = pooled if pool else z
        x, xl, cin = z, zl, cout
    return (out, xl) if want_lo else out


class _MLP(Function):
    """x_cl [rows, pad4(cin)] -> relu(bn(conv(.))) stack -> [rows, pad4(cout)]  (or [rows/pool_u, pad4(cout)])."""

    @staticmethod
    def forward(ctx, x_cl, x_lo, group_bias, meta, *params):
        lib = _lib_sizes()
        dev = x_cl.device
        rows = x_cl.shape[0]
        npass, training, pool_u = meta["npass"], meta["training"], meta["pool_u"]
        widths, bns = meta["widths"], meta["bns"]
        cin = meta["cin"]
        saved = []
        x, xl = x_cl, (x_lo if npass > 1 else None)
        out = None
        nl = len(widths)
        for li, cout in enumerate(widths):
            w, bias, gamma, beta = params[4 * li:4 * li + 4]
            bn = bns[li]
            co = _pad4(cout)
            last = li == nl - 1
            pool = pool_u if (last and pool_u) else 0
            keep = meta["need_bwd"]
            alloc = (lambda *s, dtype=torch.float32: torch.empty(s, dtype=dtype, device=dev))
            y = alloc(rows, co) if keep else _scratch("mlp_y%d" % (li & 1), rows * co, dev).view(-1)[:rows * co].view(rows, co)
            coef = alloc(4 * co)
            wprep = _scratch("mlp_wprep", lib.pvcnn_mlp_wprep_floats(cin, cout), dev)
            partials = _scratch("mlp_partials", lib.pvcnn_mlp_partials_floats(cout), dev)
            z = zl = pooled = argmax = tmp = None
            if pool:
                groups = rows // pool
                pooled = alloc(groups, co)
                argmax = alloc(groups, co, dtype=torch.int32)
                segs = lib.pvcnn_mlp_pool_segments(_LL(groups), pool)
                if segs > 1:
                    tmp = _scratch("mlp_pooltmp", 2 * groups * segs * co, dev)
            else:
                z = alloc(rows, co)
                zl = alloc(rows, co) if (npass > 1 and not last) else None
            _lib.call("pvcnn_mlp_layer_forward", _LL(rows), cin, cout, int(training), npass, float(bn.eps),
                      float(bn.momentum if bn.momentum is not None else 0.1), x, xl, w.detach(),
                      None if bias is None else bias.detach(), gamma.detach(), beta.detach(),
                      bn.running_mean if (bn.track_running_stats and bn.running_mean is not None) else None,
                      bn.running_var if (bn.track_running_stats and bn.running_var is not None) else None,
                      bn.num_batches_tracked if (training and bn.num_batches_tracked is not None
                                                 and bn.num_batches_tracked.dtype == torch.int64) else None,
                      wprep, partials, coef, y, z, zl, pool, pooled, argmax, tmp,
                      _LL(meta["group_rows"] if (li == 0 and group_bias is not None) else 0),
                      group_bias if li == 0 else None, 0 if group_bias is None else group_bias.shape[1], device=dev)
            if keep:
                saved.append((x, xl, y, coef, argmax))
            out = pooled if pool else z
            x, xl, cin = z, zl, cout
        ctx.saved = saved
        ctx.has_group_bias = group_bias is not None
        ctx.meta = dict(meta, rows=rows)
        ctx.params = [None if p is None else p.detach() for p in params]
        return out

    @staticmethod
    def backward(ctx, g):
        meta = ctx.meta
        if not meta["need_bwd"]:
            raise RuntimeError("SharedMLP forward ran without saving activations (no_grad / eval); cannot run backward")
        lib = _lib_sizes()
        dev = g.device
        rows, npass, pool_u, widths = meta["rows"], meta["npass"], meta["pool_u"], meta["widths"]
        cins = [meta["cin"]] + list(widths[:-1])
        grads = [None] * len(ctx.params)
        g = g.contiguous().float()
        nl = len(widths)
        d_gb = None
        for li in range(nl - 1, -1, -1):
            cin, cout = cins[li], widths[li]
            ci, co = _pad4(cin), _pad4(cout)
            x, xl, y, coef, argmax = ctx.saved[li]
            w = ctx.params[4 * li]
            if li == nl - 1 and pool_u:
                gz = _scratch("mlp_gz_pool", rows * co, dev)
                _lib.call("pvcnn_mlp_pool_backward", _LL(rows // pool_u), pool_u, cout, g, argmax, gz, device=dev)
                g = gz
            gy = _scratch("mlp_gy", rows * co, dev)
            gyl = _scratch("mlp_gy_lo", rows * co, dev) if npass > 1 else None
            need_gx = li > 0 or meta["need_input_grad"]
            gx = torch.empty((rows, ci), dtype=torch.float32, device=dev) if need_gx else None
            dw = torch.empty_like(w)
            dbias = torch.empty(cout, dtype=torch.float32, device=dev)
            dgamma = torch.empty(cout, dtype=torch.float32, device=dev)
            dbeta = torch.empty(cout, dtype=torch.float32, device=dev)
            wprep = _scratch("mlp_wprep", lib.pvcnn_mlp_wprep_floats(cin, cout), dev)
            partials = _scratch("mlp_partials", lib.pvcnn_mlp_partials_floats(cout), dev)
            sums = _scratch("mlp_sums", 4 * co, dev)
            if li == 0 and ctx.has_group_bias and ctx.needs_input_grad[2]:
                d_gb = torch.empty((rows // meta["group_rows"], co), dtype=torch.float32, device=dev)
            _lib.call("pvcnn_mlp_layer_backward", _LL(rows), cin, cout, npass, g, x, xl, w, y, coef, wprep, partials,
                      sums, gy, gyl, gx, dw, dbias, dgamma, dbeta, _LL(meta["group_rows"] if li == 0 else 0),
                      d_gb if li == 0 else None, device=dev)
            grads[4 * li] = dw
            grads[4 * li + 1] = dbias if ctx.params[4 * li + 1] is not None else None
            grads[4 * li + 2] = dgamma
            grads[4 * li + 3] = dbeta
            g = gx
        return (g, None, d_gb, None, *grads)


def native_supported(layers):
    """True when `layers` is the reference's (conv k=1, BatchNorm, ReLU)* stack with what the kernels assume."""
    mods = list(layers)
    if len(mods) == 0 or len(mods) % 3 != 0:
        return False
    for i in range(0, len(mods), 3):
        conv, bn, act = mods[i:i + 3]
        if not isinstance(conv, (torch.nn.Conv1d, torch.nn.Conv2d)) or not isinstance(bn, torch.nn.modules.batchnorm._BatchNorm):
            return False
        if any(k != 1 for k in conv.kernel_size) or any(s != 1 for s in conv.stride) or conv.groups != 1:
            return False
        if any(p != 0 for p in conv.padding) or conv.padding_mode != "zeros":
            return False
        if not bn.affine or not bn.track_running_stats or bn.momentum is None or not isinstance(act, torch.nn.ReLU):
            return False
        if conv.out_channels > 1024 or conv.weight.dtype != torch.float32:
            return False
    return True


def mlp_cl(layers, x_cl, x_lo, pool_u=0, input_needs_grad=True, point_cols=None, group_bias=None, group_rows=0,
           want_lo=False):
    """Run the (conv, bn, relu)* stack `layers` on channels-last rows.  Returns [rows, pad4(cout)] (or pooled).

    point_cols / group_bias / group_rows: the first layer's input is the reference's concatenation restricted to the
    channel ranges `point_cols`; the remaining channels are constant over each cloud of `group_rows` rows and enter as
    group_bias [clouds, pad4(cout)] = their part of the weight applied once per cloud (see head_cl).
    want_lo: return (out, out_lo | None); the lo operand of the next 3xTF32 GEMM comes for free from the inference epilogue."""
    mods = list(layers)
    convs, bns = mods[0::3], mods[1::3]
    training = bool(bns[0].training)
    fused_eval = (not training and not torch.is_grad_enabled()
                  and os.environ.get("PVCNN_B200_MLP_EVAL", "fused") != "layers")
    params = []
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        w = conv.weight
        if i == 0 and point_cols is not None and not fused_eval:
            w = _select_cols(w, point_cols)          # autograd: the gradient flows back into the full weight
        params += [w, conv.bias, bn.weight, bn.bias]
    need_bwd = training and torch.is_grad_enabled() and (
        (input_needs_grad and x_cl.requires_grad) or any(p is not None and p.requires_grad for p in params)
        or (group_bias is not None and group_bias.requires_grad))
    cin = convs[0].in_channels if point_cols is None else sum(b - a for a, b in point_cols)
    meta = dict(npass=precision_passes(), training=training, pool_u=int(pool_u), widths=[c.out_channels for c in convs],
                bns=bns, convs=convs, cin=cin, need_bwd=bool(need_bwd), group_rows=int(group_rows),
                point_cols=None if point_cols is None else tuple(point_cols),
                need_input_grad=bool(input_needs_grad and x_cl.requires_grad))
    if fused_eval:
        # inference: frozen-layer path (cached weight operands / BatchNorm coefficients, activation fused into the GEMM)
        return _mlp_eval(x_cl.detach(), None if x_lo is None else x_lo.detach(), meta,
                         None if group_bias is None else group_bias.detach(), group_rows, want_lo=want_lo)
    out = _run(_MLP, x_cl, x_lo, group_bias, meta, *params)
    return (out, None) if want_lo else out


def shared_mlp_forward(layers, x):
    """x [B, C, *spatial] -> [B, Cout, *spatial] through the native path."""
    b, c = x.shape[:2]
    spatial = tuple(x.shape[2:])
    n = 1
    for s in spatial:
        n *= s
    x_cl, x_lo = _run(_ToCL, x.reshape(b, c, n), precision_passes() > 1)
    z = mlp_cl(layers, x_cl, x_lo if x_lo.numel() else None)
    cout = list(layers)[-3].out_channels
    return _run(_FromCL, z, b, cout, n).view(b, cout, *spatial)


class _GroupConcatCL(Function):
    """modules/ball_query.py:16-30 with channels-last output rows (b, m, u) x pad4(3 + C)  (+ lo)."""

    @staticmethod
    def forward(ctx, points_coords, centers_coords, features, indices, want_lo):
        points_coords = points_coords.contiguous().float()
        centers_coords = centers_coords.contiguous().float()
        features = None if features is None else features.contiguous().float()
        indices = indices.int().contiguous()
        b, _, n = points_coords.shape
        _, m, u = indices.shape
        c = 0 if features is None else features.shape[1]
        cp = _pad4(3 + c)
        out = torch.empty((b * m * u, cp), dtype=torch.float32, device=points_coords.device)
        lo = torch.empty_like(out) if want_lo else None
        _lib.call("pvcnn_group_concat_cl", b, c, n, m, u, points_coords, centers_coords, features, indices, out, lo)
        ctx.save_for_backward(indices)
        ctx.dims = (b, c, n, m, u)
        if lo is None:
            lo = out.new_empty(0)
        ctx.mark_non_differentiable(lo)
        return out, lo

    @staticmethod
    def backward(ctx, g, _glo):
        (indices,) = ctx.saved_tensors
        b, c, n, m, u = ctx.dims
        dev = g.device
        need_p, need_c, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and c > 0
        gf = torch.empty((b, c, n), dtype=torch.float32, device=dev) if need_f else None
        gp = torch.empty((b, 3, n), dtype=torch.float32, device=dev) if need_p else None
        gc = torch.empty((b, 3, m), dtype=torch.float32, device=dev) if need_c else None
        if need_f or need_p or need_c:
            _lib.call("pvcnn_group_concat_cl_grad", b, c, n, m, u, g.contiguous(), indices, gf, gp, gc)
        return gp, gc, gf, None, None


def sa_branch(layers, coords, centers, features, indices):
    """PointNet++ set-abstraction branch (modules/pointnet.py:85-87) on the native path:
    grouping -> channels-last rows -> tensor-core MLP -> max over the U neighbours -> [B, Cout, M]."""
    b = coords.shape[0]
    _, m, u = indices.shape
    rows, lo = _run(_GroupConcatCL, coords, centers, features, indices, precision_passes() > 1)
    pooled = mlp_cl(layers, rows, lo if lo.numel() else None, pool_u=u)
    cout = list(layers)[-3].out_channels
    return _run(_FromCL, pooled, b, cout, m)


class _CatCL(Function):
    """torch.cat(tensors, dim=1) of [B,C_i,N] (or [B,C_i,1], broadcast over N) written straight into channels-last rows
    [B*N, pad4(sum C_i)] (+ lo).  Model-level glue of models/s3dis/pvcnn.py:44-46 / models/shapenet/pvcnn.py:40-42."""

    @staticmethod
    def forward(ctx, n, want_lo, *tensors):
        b = tensors[0].shape[0]
        dev = tensors[0].device
        widths = [t.shape[1] for t in tensors]
        ctot = sum(widths)
        ld = _pad4(ctot)
        rows = torch.empty((b * n, ld), dtype=torch.float32, device=dev)
        lo = torch.empty_like(rows) if want_lo else None
        if ld != ctot:
            rows[:, ctot:].zero_()
            if lo is not None:
                lo[:, ctot:].zero_()
        col = 0
        for t in tensors:
            t = t.contiguous().float()
            _lib.call("pvcnn_cat_to_cl", b, t.shape[1], n, t.shape[2], t, ld, col, rows, lo, device=dev)
            col += t.shape[1]
        ctx.meta = (b, n, ld, widths, [t.shape[2] for t in tensors])
        if lo is None:
            lo = rows.new_empty(0)
        ctx.mark_non_differentiable(lo)
        return rows, lo

    @staticmethod
    def backward(ctx, g, _glo):
        b, n, ld, widths, src_ns = ctx.meta
        g = g.contiguous()
        outs = []
        col = 0
        for i, (c, sn) in enumerate(zip(widths, src_ns)):
            if ctx.needs_input_grad[2 + i]:
                gx = torch.empty((b, c, n), dtype=torch.float32, device=g.device)
                _lib.call("pvcnn_cl_slice_to_points", b, c, n, g, ld, col, gx, device=g.device)
                outs.append(gx.sum(dim=2, keepdim=True) if sn == 1 else gx)
            else:
                outs.append(None)
            col += c
        return (None, None, *outs)


def cat_cl(tensors, n):
    """-> (rows [B*N, pad4(sum C)], lo | None)"""
    rows, lo = _run(_CatCL, n, precision_passes() > 1, *tensors)
    return rows, (lo if lo.numel() else None)


class _LinearCL(Function):
    """Conv1d(k=1) without BatchNorm / ReLU on channels-last rows (the classifier's last layer)."""

    @staticmethod
    def forward(ctx, x, x_lo, w, bias):
        lib = _lib_sizes()
        dev = x.device
        rows = x.shape[0]
        cout, cin = w.shape[0], w.shape[1]
        npass = precision_passes()
        if npass > 1 and x_lo is None:
            from . import dense
            x_lo = dense.split_tf32(x, want_hi=False)[1]
        y = torch.empty((rows, _pad4(cout)), dtype=torch.float32, device=dev)
        wprep = _scratch("mlp_wprep", lib.pvcnn_mlp_wprep_floats(cin, cout), dev)
        _lib.call("pvcnn_linear_cl_forward", _LL(rows), cin, cout, npass, x, x_lo, w.detach(),
                  None if bias is None else bias.detach(), wprep, y, device=dev)
        ctx.save_for_backward(x, x_lo, w.detach())
        ctx.meta = (rows, cin, cout, npass, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib_sizes()
        x, x_lo, w = ctx.saved_tensors
        rows, cin, cout, npass, has_bias = ctx.meta
        dev = gy.device
        gy = gy.contiguous()
        gy_lo = None
        if npass > 1:
            from . import dense
            gy_lo = dense.split_tf32(gy, want_hi=False)[1]
        need_gx = ctx.needs_input_grad[0]
        gx = torch.empty((rows, _pad4(cin)), dtype=torch.float32, device=dev) if need_gx else None
        dw = torch.empty_like(w)
        db = torch.empty(cout, dtype=torch.float32, device=dev) if has_bias else None
        wprep = _scratch("mlp_wprep", lib.pvcnn_mlp_wprep_floats(cin, cout), dev)
        partials = _scratch("mlp_partials", lib.pvcnn_mlp_partials_floats(cout), dev)
        _lib.call("pvcnn_linear_cl_backward", _LL(rows), cin, cout, npass, gy, gy_lo, x, x_lo, w, wprep, partials, gx, dw,
                  db, device=dev)
        return gx, None, dw, db


def head_supported(seq):
    from .nn.shared_mlp import SharedMLP
    for m in seq:
        ok = isinstance(m, torch.nn.Dropout) or (isinstance(m, SharedMLP) and native_supported(m.layers)) or (
            isinstance(m, torch.nn.Conv1d) and m.kernel_size == (1,) and m.groups == 1 and m.stride == (1,))
        if not ok:
            return False
    return True


def cloud_bias(conv, cloud_taps, cloud_cols):
    """The first head layer's response to the channels that are constant over a cloud: [B, pad4(cout)] =
    W[:, cloud_cols] @ cat(cloud_taps), one small GEMM per forward instead of K_cloud extra columns on every point row."""
    g = torch.cat([t.reshape(t.shape[0], t.shape[1]) for t in cloud_taps], dim=1).float()
    kb = g.shape[1]
    if _pad4(kb) != kb:
        g = torch.nn.functional.pad(g, (0, _pad4(kb) - kb))
    w_b = _select_cols(conv.weight, cloud_cols)
    return _run(_LinearCL, g.contiguous(), None, w_b, None)


def head_cl(seq, rows, lo, b, n, point_cols=None, group_bias=None):
    """Run a per-point head -- nn.Sequential of SharedMLP | nn.Dropout | nn.Conv1d(k=1) (models/utils.py:15-45) -- on
    channels-last rows and return [B, Cout, N].  Returns None if the head holds anything else (caller falls back).
    point_cols / group_bias: see mlp_cl (they apply to the first module, which must then be a SharedMLP)."""
    from .nn.shared_mlp import SharedMLP
    mods = list(seq)
    for m in mods:
        ok = isinstance(m, torch.nn.Dropout) or (isinstance(m, SharedMLP) and native_supported(m.layers)) or (
            isinstance(m, torch.nn.Conv1d) and m.kernel_size == (1,) and m.groups == 1 and m.stride == (1,))
        if not ok:
            return None
    x, xl, cout = rows, lo, None
    for m in mods:
        if isinstance(m, torch.nn.Dropout):
            if m.training and m.p > 0:
                x = torch.nn.functional.dropout(x, m.p, True)   # element-wise i.i.d.: layout-agnostic
                xl = None
        elif isinstance(m, SharedMLP):
            if xl is None and precision_passes() > 1:
                from . import dense
                xl = dense.split_tf32(x.contiguous(), want_hi=False)[1]
            if m is mods[0] and group_bias is not None:
                x, xl = mlp_cl(m.layers, x, xl, point_cols=point_cols, group_bias=group_bias, group_rows=n, want_lo=True)
            else:
                x, xl = mlp_cl(m.layers, x, xl, want_lo=True)
            cout = list(m.layers)[-3].out_channels
        else:
            x = _run(_LinearCL, x, xl, m.weight, m.bias)
            xl = None
            cout = m.out_channels
    return _run(_FromCL, x, b, cout, n)
