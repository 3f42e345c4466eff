"""numpy-facing wrapper around oracle/pvcnn_oracle.c plus the torch-CPU restatement of the
dense (third-party) arithmetic of PVConv.

Parity status: the reference ships no tests or golden vectors, so the oracle is
pinned against the reference's *own CUDA kernels*, compiled unmodified into oracle/_ref/ and run
on a B200: golden outputs in tests/golden/ref_ops_golden.npz (tests/test_golden_cpu.py), live comparison in
tests/test_ops_gpu.py.  The dense ops (Conv3d / BatchNorm / Conv1d, which
the reference delegates to torch, modules/pvconv.py:21-26, modules/shared_mlp.py:10-24) are
restated with the same torch calls on CPU in fp32 / fp64.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpvcnn_oracle.so")

__all__ = [
    "build", "num_threads", "voxelize_coords", "avg_voxelize", "avg_voxelize_grad",
    "trilinear_devoxelize", "trilinear_devoxelize_grad", "ball_query", "grouping", "grouping_grad", "group_concat", "group_concat_grad",
    "gather", "gather_grad", "furthest_point_sampling", "three_nn", "three_nn_interpolate",
    "three_nn_interpolate_grad", "pvconv_forward_backward", "logits_mask_sample", "torch_mean",
]


def build(force=False):
    src = os.path.join(_HERE, "pvcnn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_ci = ctypes.c_int
_cf = ctypes.c_float


def torch_mean(coords):
    """The reference's own mean: `coords.mean(2, keepdim=True)` (modules/voxelization.py:18) as torch computes it on
    CPU.  GPU tests pass the same call's result from the device under test instead (mean=...)."""
    import torch
    return torch.from_numpy(_f(coords)).mean(2, keepdim=True).numpy().reshape(-1, 3)


def voxelize_coords(coords, r, normalize=True, eps=0.0, mean=None):
    """modules/voxelization.py:17-24.  mean: [B,3] result of the reference's `coords.mean(2)` (default: torch CPU);
    mean="fp64" selects the round-1 definition (fp64 sum rounded once)."""
    coords = _f(coords)
    b, _, n = coords.shape
    nc = np.empty_like(coords)
    vc = np.empty(coords.shape, dtype=np.int32)
    if isinstance(mean, str) and mean == "fp64":
        lib().oracle_voxelize_coords(_ci(b), _ci(n), _ci(r), _ci(int(normalize)), _cf(eps), _p(coords), _p(nc), _p(vc))
        return nc, vc
    mean = _f(torch_mean(coords) if mean is None else np.asarray(mean).reshape(-1, 3))
    assert mean.shape == (b, 3)
    lib().oracle_voxelize_coords_given(_ci(b), _ci(n), _ci(r), _ci(int(normalize)), _cf(eps), _p(coords), _p(mean),
                                       _p(nc), _p(vc))
    return nc, vc


def avg_voxelize(feat, coords, r):
    feat, coords = _f(feat), _i(coords)
    b, c, n = feat.shape
    ind = np.empty((b, n), np.int32)
    cnt = np.empty((b, r ** 3), np.int32)
    out = np.empty((b, c, r ** 3), np.float32)
    lib().oracle_avg_voxelize(_ci(b), _ci(c), _ci(n), _ci(r), _p(coords), _p(feat), _p(ind), _p(cnt), _p(out))
    return out, ind, cnt


def avg_voxelize_grad(grad_y, ind, cnt):
    grad_y, ind, cnt = _f(grad_y), _i(ind), _i(cnt)
    b, c, r3 = grad_y.shape
    n = ind.shape[1]
    gx = np.empty((b, c, n), np.float32)
    lib().oracle_avg_voxelize_grad(_ci(b), _ci(c), _ci(n), _ci(r3), _p(ind), _p(cnt), _p(grad_y), _p(gx))
    return gx


def trilinear_devoxelize(feat, coords, r, is_training=True):
    feat, coords = _f(feat), _f(coords)
    b, c = feat.shape[:2]
    feat = feat.reshape(b, c, -1)
    n = coords.shape[2]
    inds = np.zeros((b, 8, n), np.int32)
    wgts = np.zeros((b, 8, n), np.float32)
    outs = np.empty((b, c, n), np.float32)
    lib().oracle_trilinear_devoxelize(_ci(b), _ci(c), _ci(n), _ci(r), _ci(int(is_training)), _p(coords), _p(feat),
                                      _p(inds), _p(wgts), _p(outs))
    return outs, inds, wgts


def trilinear_devoxelize_grad(grad_y, inds, wgts, r):
    grad_y, inds, wgts = _f(grad_y), _i(inds), _f(wgts)
    b, c, n = grad_y.shape
    gx = np.empty((b, c, r ** 3), np.float32)
    lib().oracle_trilinear_devoxelize_grad(_ci(b), _ci(c), _ci(n), _ci(r ** 3), _p(inds), _p(wgts), _p(grad_y), _p(gx))
    return gx


def ball_query(centers, points, radius, u):
    centers, points = _f(centers), _f(points)
    b, _, m = centers.shape
    n = points.shape[2]
    out = np.empty((b, m, u), np.int32)
    r2 = np.float32(radius) * np.float32(radius)  # ball_query.cpp:24 (float product)
    lib().oracle_ball_query(_ci(b), _ci(n), _ci(m), _cf(r2), _ci(u), _p(centers), _p(points), _p(out))
    return out


def grouping(feat, idx):
    feat, idx = _f(feat), _i(idx)
    b, c, n = feat.shape
    _, m, u = idx.shape
    out = np.empty((b, c, m, u), np.float32)
    lib().oracle_grouping(_ci(b), _ci(c), _ci(n), _ci(m), _ci(u), _p(feat), _p(idx), _p(out))
    return out


def grouping_grad(grad_y, idx, n):
    grad_y, idx = _f(grad_y), _i(idx)
    b, c, m, u = grad_y.shape
    gx = np.empty((b, c, n), np.float32)
    lib().oracle_grouping_grad(_ci(b), _ci(c), _ci(n), _ci(m), _ci(u), _p(grad_y), _p(idx), _p(gx))
    return gx


def group_concat(points_coords, centers_coords, feat, idx):
    """modules/ball_query.py:16-30 restated on the oracle's grouping: cat([grouping(coords) - centres, grouping(feat)])."""
    rel = grouping(points_coords, idx) - _f(centers_coords)[:, :, :, None]
    return rel if feat is None else np.concatenate([rel, grouping(feat, idx)], axis=1)


def group_concat_grad(grad_y, idx, n):
    """-> (grad_features [B,C,N] | None, grad_points_coords [B,3,N], grad_centers_coords [B,3,M])"""
    grad_y = _f(grad_y)
    gxyz = np.ascontiguousarray(grad_y[:, :3])
    gf = grouping_grad(np.ascontiguousarray(grad_y[:, 3:]), idx, n) if grad_y.shape[1] > 3 else None
    return gf, grouping_grad(gxyz, idx, n), -gxyz.sum(axis=3, dtype=np.float64).astype(np.float32)


def gather(feat, idx):
    feat, idx = _f(feat), _i(idx)
    b, c, n = feat.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().oracle_gather(_ci(b), _ci(c), _ci(n), _ci(m), _p(feat), _p(idx), _p(out))
    return out


def gather_grad(grad_y, idx, n):
    grad_y, idx = _f(grad_y), _i(idx)
    b, c, m = grad_y.shape
    gx = np.empty((b, c, n), np.float32)
    lib().oracle_gather_grad(_ci(b), _ci(c), _ci(n), _ci(m), _p(grad_y), _p(idx), _p(gx))
    return gx


def furthest_point_sampling(coords, m):
    coords = _f(coords)
    b, _, n = coords.shape
    out = np.zeros((b, m), np.int32)
    lib().oracle_furthest_point_sampling(_ci(b), _ci(n), _ci(m), _p(coords), _p(out))
    return out


def three_nn(points, centers):
    points, centers = _f(points), _f(centers)
    b, _, n = points.shape
    m = centers.shape[2]
    w = np.empty((b, 3, n), np.float32)
    idx = np.empty((b, 3, n), np.int32)
    lib().oracle_three_nn(_ci(b), _ci(n), _ci(m), _p(points), _p(centers), _p(w), _p(idx))
    return idx, w


def three_nn_interpolate(cfeat, idx, w):
    cfeat, idx, w = _f(cfeat), _i(idx), _f(w)
    b, c, m = cfeat.shape
    n = idx.shape[2]
    out = np.empty((b, c, n), np.float32)
    lib().oracle_three_nn_interpolate(_ci(b), _ci(c), _ci(m), _ci(n), _p(cfeat), _p(idx), _p(w), _p(out))
    return out


def three_nn_interpolate_grad(grad_y, idx, w, m):
    grad_y, idx, w = _f(grad_y), _i(idx), _f(w)
    b, c, n = grad_y.shape
    gx = np.empty((b, c, m), np.float32)
    lib().oracle_three_nn_interpolate_grad(_ci(b), _ci(c), _ci(n), _ci(m), _p(grad_y), _p(idx), _p(w), _p(gx))
    return gx


def _lm_mix(seed, b, stream, j):
    """splitmix64 finaliser of (seed, sample, stream, index) -- the counter-based generator of pvcnn_logits_mask_sample"""
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) ^ (np.uint64(b) * np.uint64(0x9E3779B97F4A7C15)) ^ (np.uint64(stream) * np.uint64(0xBF58476D1CE4E5B9))
             ^ (np.asarray(j, dtype=np.uint64) * np.uint64(0x94D049BB133111EB)))
        x = x ^ (x >> np.uint64(30)); x = x * np.uint64(0xBF58476D1CE4E5B9)
        x = x ^ (x >> np.uint64(27)); x = x * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def logits_mask_sample(mask, k, seed):
    """Restatement of the resampling of modules/functional/sampling.py:66-82 with the device generator instead of numpy's:
    nc >= k: the k candidates with the smallest random keys (a uniform k-subset in random order, = np.random.choice(nc, k,
    replace=False)); 0 < nc < k: arange(nc).repeat(k // nc) ++ (k % nc)-subset, shuffled by a second key stream; nc == 0:
    zeros.  mask: bool [B,N] -> int32 [B,k]."""
    mask = np.asarray(mask).astype(bool)
    bsz, n = mask.shape
    out = np.zeros((bsz, k), np.int32)
    hi = np.uint64(0xFFFFFFFF00000000)
    for b in range(bsz):
        cand = np.nonzero(mask[b])[0].astype(np.int32)
        nc = cand.size
        if nc == 0:
            continue
        j = np.arange(nc, dtype=np.uint64)
        order = np.argsort((_lm_mix(seed, b, 0, j) & hi) | j, kind="stable")
        if nc >= k:
            out[b] = cand[order[:k]]
            continue
        rep, rem = k // nc, k % nc
        lst = np.concatenate([np.arange(nc).repeat(rep), order[:rem]]).astype(np.int64)
        t = np.arange(k, dtype=np.uint64)
        perm = np.argsort((_lm_mix(seed, b, 1, t) & hi) | t, kind="stable")
        out[b] = cand[lst[perm]]
    return out


# ----------------------------------------------------------------------------------------------
# PVConv block oracle: modules/pvconv.py:33-39 wiring, dense ops through torch CPU.
# ----------------------------------------------------------------------------------------------
def pvconv_forward_backward(params, features, coords, grad_out, resolution, *, training=True, normalize=True,
                            eps=0.0, with_se=False, dtype="float32", bn_eps=1e-4, momentum=0.1, buffers=None, threads=None,
                            vox_mean=None):
    """Forward (+ backward when grad_out is not None) of one PVConv block on CPU.

    params: dict with the reference state_dict names:
      voxel_layers.{0,3}.{weight,bias}, voxel_layers.{1,4}.{weight,bias}, point_features.layers.0.{weight,bias},
      point_features.layers.1.{weight,bias} [, voxel_layers.6.fc.{0,2}.weight]
    buffers: optional dict of running stats (used when training=False).
    Returns dict(out=..., grads={name: ...}, grad_features=..., stats={...}).
    The custom ops go through the C oracle in fp32 (they are fp32 in the reference); the dense ops
    run in `dtype` (float32 mirrors the reference with allow_tf32=False, float64 is the truth).
    """
    import torch
    import torch.nn.functional as TF

    if threads is not None:
        torch.set_num_threads(int(threads))
    td = getattr(torch, dtype)
    r = int(resolution)
    feats = torch.as_tensor(np.asarray(features), dtype=torch.float32)
    P = {k: torch.as_tensor(np.asarray(v)).to(td).requires_grad_(grad_out is not None) for k, v in params.items()}
    b, c_in, n = feats.shape

    class _Vox(torch.autograd.Function):
        @staticmethod
        def forward(ctx, f, vc):
            out, ind, cnt = avg_voxelize(f.detach().float().numpy(), vc, r)
            ctx.ind, ctx.cnt = ind, cnt
            return torch.from_numpy(out).to(f.dtype).view(f.shape[0], f.shape[1], r, r, r)

        @staticmethod
        def backward(ctx, g):
            gy = g.contiguous().view(g.shape[0], g.shape[1], -1).float().numpy()
            return torch.from_numpy(avg_voxelize_grad(gy, ctx.ind, ctx.cnt)).to(g.dtype), None

    class _Devox(torch.autograd.Function):
        @staticmethod
        def forward(ctx, f, nc):
            out, inds, wgts = trilinear_devoxelize(f.detach().float().numpy(), nc, r, True)
            ctx.inds, ctx.wgts = inds, wgts
            return torch.from_numpy(out).to(f.dtype)

        @staticmethod
        def backward(ctx, g):
            gx = trilinear_devoxelize_grad(g.contiguous().float().numpy(), ctx.inds, ctx.wgts, r)
            return torch.from_numpy(gx).to(g.dtype).view(g.shape[0], g.shape[1], r, r, r), None

    class _VoxD(torch.autograd.Function):
        """fp64 'truth' variants keep the scatter/gather in the working dtype."""
        @staticmethod
        def forward(ctx, f, ind_t, cnt_t):
            bsz, ch, _ = f.shape
            inv = torch.where(cnt_t > 0, 1.0 / cnt_t.clamp(min=1).to(f.dtype), torch.zeros((), dtype=f.dtype))
            w = torch.gather(inv, 1, ind_t)  # [b, n]
            out = torch.zeros(bsz, ch, r ** 3, dtype=f.dtype)
            out.scatter_add_(2, ind_t.unsqueeze(1).expand(-1, ch, -1), f * w.unsqueeze(1))
            ctx.save_for_backward(ind_t, w)
            return out.view(bsz, ch, r, r, r)

        @staticmethod
        def backward(ctx, g):
            ind_t, w = ctx.saved_tensors
            bsz, ch = g.shape[:2]
            gg = torch.gather(g.reshape(bsz, ch, -1), 2, ind_t.unsqueeze(1).expand(-1, ch, -1))
            return gg * w.unsqueeze(1), None, None

    x = feats.to(td).requires_grad_(grad_out is not None)
    nc, vc = voxelize_coords(np.asarray(coords, dtype=np.float32), r, normalize, eps, mean=vox_mean)

    if td == torch.float32:
        vox = _Vox.apply(x, vc)
    else:
        _, ind, cnt = avg_voxelize(np.zeros((b, 1, n), np.float32), vc, r)
        vox = _VoxD.apply(x, torch.from_numpy(ind).long(), torch.from_numpy(cnt).long())

    def bn(t, prefix):
        rm = rv = None
        if buffers is not None:
            rm = torch.as_tensor(np.asarray(buffers[prefix + ".running_mean"])).to(td).clone()
            rv = torch.as_tensor(np.asarray(buffers[prefix + ".running_var"])).to(td).clone()
        e = bn_eps if prefix.startswith("voxel") else 1e-5
        y = TF.batch_norm(t, rm, rv, P[prefix + ".weight"], P[prefix + ".bias"], training or rm is None, momentum, e)
        return y, rm, rv

    stats = {}
    h = TF.conv3d(vox, P["voxel_layers.0.weight"], P["voxel_layers.0.bias"], padding=1)
    stats["conv1"] = h.detach()
    h, rm, rv = bn(h, "voxel_layers.1"); stats["bn1_running"] = (rm, rv)
    h = TF.leaky_relu(h, 0.1)
    h = TF.conv3d(h, P["voxel_layers.3.weight"], P["voxel_layers.3.bias"], padding=1)
    stats["conv2"] = h.detach()
    h, rm, rv = bn(h, "voxel_layers.4"); stats["bn2_running"] = (rm, rv)
    h = TF.leaky_relu(h, 0.1)
    if with_se:
        s = h.mean(-1).mean(-1).mean(-1)
        s = torch.sigmoid(TF.linear(torch.relu(TF.linear(s, P["voxel_layers.6.fc.0.weight"])),
                                    P["voxel_layers.6.fc.2.weight"]))
        h = h * s.view(b, -1, 1, 1, 1)
    if td == torch.float32:
        vfeat = _Devox.apply(h, nc)
    else:
        _, inds, wgts = trilinear_devoxelize(np.zeros((b, 1, r ** 3), np.float32), nc, r, True)
        it = torch.from_numpy(inds).long()
        wt = torch.from_numpy(wgts).to(td)
        hv = h.reshape(b, h.shape[1], -1)
        vfeat = 0
        for k in range(8):
            vfeat = vfeat + torch.gather(hv, 2, it[:, k:k + 1, :].expand(-1, hv.shape[1], -1)) * wt[:, k:k + 1, :]
    p = TF.conv1d(x, P["point_features.layers.0.weight"], P["point_features.layers.0.bias"])
    p, rm, rv = bn(p, "point_features.layers.1"); stats["bnp_running"] = (rm, rv)
    p = torch.relu(p)
    out = vfeat + p
    res = {"out": out.detach().numpy(), "norm_coords": nc, "vox_coords": vc, "stats": stats}
    if grad_out is not None:
        go = torch.as_tensor(np.asarray(grad_out)).to(td)
        out.backward(go)
        res["grad_features"] = x.grad.numpy()
        res["grads"] = {k: v.grad.numpy() for k, v in P.items() if v.grad is not None}
    return res
