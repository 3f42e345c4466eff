"""autograd.Function wrappers -- same contracts as the reference's
modules/functional/{voxelization,devoxelization,ball_query,grouping,sampling,interpolatation}.py."""
import ctypes

import numpy as np
import os

import torch
from torch.autograd import Function

from . import backend as _backend
from .. import _lib


class _AvgVoxelize(Function):
    """modules/functional/voxelization.py:8-37: features [B,C,N], integer coords [B,3,N] ->
    [B,C,R,R,R]; saves (ind, cnt); gradient flows to features only."""

    @staticmethod
    def forward(ctx, features, coords, resolution):
        features = features.contiguous()
        coords = coords.int().contiguous()
        b, c, _ = features.shape
        out, ind, cnt = _backend.avg_voxelize_forward(features, coords, resolution)
        ctx.save_for_backward(ind, cnt)
        return out.view(b, c, resolution, resolution, resolution)

    @staticmethod
    def backward(ctx, grad_output):
        ind, cnt = ctx.saved_tensors
        b, c = grad_output.shape[:2]
        g = _backend.avg_voxelize_backward(grad_output.contiguous().view(b, c, -1), ind, cnt)
        return g, None, None


avg_voxelize = _AvgVoxelize.apply


class _TrilinearDevoxelize(Function):
    """modules/functional/devoxelization.py:8-39: grid [B,C,R,R,R] sampled at float coords
    [B,3,N] -> [B,C,N]; (inds, wgts) are kept only in training mode."""

    @staticmethod
    def forward(ctx, features, coords, resolution, is_training=True):
        b, c = features.shape[:2]
        features = features.contiguous().view(b, c, -1)
        coords = coords.contiguous()
        outs, inds, wgts = _backend.trilinear_devoxelize_forward(resolution, is_training, coords, features)
        if is_training:
            ctx.save_for_backward(inds, wgts)
            ctx.r = resolution
        return outs

    @staticmethod
    def backward(ctx, grad_output):
        inds, wgts = ctx.saved_tensors
        g = _backend.trilinear_devoxelize_backward(grad_output.contiguous(), inds, wgts, ctx.r)
        return g.view(grad_output.size(0), grad_output.size(1), ctx.r, ctx.r, ctx.r), None, None, None


trilinear_devoxelize = _TrilinearDevoxelize.apply


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    """modules/functional/ball_query.py:8-19: int32 [B,M,U], not differentiable."""
    return _backend.ball_query(centers_coords.contiguous(), points_coords.contiguous(), radius, num_neighbors)


class _Grouping(Function):
    """modules/functional/grouping.py:8-28"""

    @staticmethod
    def forward(ctx, features, indices):
        features = features.contiguous()
        indices = indices.contiguous()
        ctx.save_for_backward(indices)
        ctx.num_points = features.size(-1)
        return _backend.grouping_forward(features, indices)

    @staticmethod
    def backward(ctx, grad_output):
        (indices,) = ctx.saved_tensors
        return _backend.grouping_backward(grad_output.contiguous(), indices, ctx.num_points), None


grouping = _Grouping.apply


class _GroupConcat(Function):
    """One-pass replacement for the tensor sequence of modules/ball_query.py:16-30
    (grouping(coords) - centres, grouping(features), cat): same values, same gradients."""

    @staticmethod
    def forward(ctx, points_coords, centers_coords, points_features, indices):
        points_coords = points_coords.contiguous()
        centers_coords = centers_coords.contiguous()
        indices = indices.contiguous()
        if points_features is not None:
            points_features = points_features.contiguous()
        ctx.save_for_backward(indices)
        ctx.num_points = points_coords.size(-1)
        ctx.has_features = points_features is not None
        return _backend.group_concat_forward(points_coords, centers_coords, points_features, indices)

    @staticmethod
    def backward(ctx, grad_output):
        (indices,) = ctx.saved_tensors
        need_p, need_c, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gf, gp, gc = _backend.group_concat_backward(grad_output.contiguous(), indices, ctx.num_points,
                                                    need_features=need_f and ctx.has_features, need_points=need_p,
                                                    need_centers=need_c)
        return gp, gc, gf, None


def group_concat(points_coords, centers_coords, points_features, indices):
    """[B,3,N], [B,3,M], [B,C,N] | None, Int[B,M,U] -> [B,3+C,M,U]"""
    return _GroupConcat.apply(points_coords, centers_coords, points_features, indices)


class _Gather(Function):
    """modules/functional/sampling.py:10-32"""

    @staticmethod
    def forward(ctx, features, indices):
        features = features.contiguous()
        indices = indices.int().contiguous()
        ctx.save_for_backward(indices)
        ctx.num_points = features.size(-1)
        return _backend.gather_features_forward(features, indices)

    @staticmethod
    def backward(ctx, grad_output):
        (indices,) = ctx.saved_tensors
        return _backend.gather_features_backward(grad_output.contiguous(), indices, ctx.num_points), None


gather = _Gather.apply


def furthest_point_sample(coords, num_samples):
    """modules/functional/sampling.py:37-48: returns the sampled *coordinates* [B,3,M]
    (differentiable through gather)."""
    coords = coords.contiguous()
    return gather(coords, _backend.furthest_point_sampling(coords, num_samples))


def logits_mask(coords, logits, num_points_per_object):
    """modules/functional/sampling.py:51-84: foreground mask from 2-way logits, masked mean, and a fixed-size
    resampling of the foreground points.

    Default: the resampling runs on the device (pvcnn_logits_mask_sample: ordered compaction + counter-based random
    subset / repeat-fill + shuffle, one CTA per sample) -- no `.nonzero()` host sync, no per-sample Python loop.  It draws
    the same distribution as the reference's numpy calls; the stream is keyed by ONE np.random draw per call, so
    np.random.seed() still makes a run reproducible.  PVCNN_B200_LOGITS_MASK=numpy keeps the reference's exact numpy
    call sequence (choice / choice+shuffle per sample, one host sync each) for bit-reproducibility against it."""
    bsz, _, npts = coords.shape
    k = int(num_points_per_object)
    mask = logits[:, 0, :] < logits[:, 1, :]
    count = mask.sum(dim=-1, keepdim=True)
    masked = coords * mask.view(bsz, 1, npts)
    mean = masked.sum(dim=-1) / torch.max(count, torch.ones_like(count)).float()
    picks = torch.zeros((bsz, k), device=coords.device, dtype=torch.int32)
    if os.environ.get("PVCNN_B200_LOGITS_MASK", "device").lower() != "numpy" and coords.is_cuda and max(npts, k) <= 8192:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
        _lib.call("pvcnn_logits_mask_sample", bsz, npts, k, ctypes.c_ulonglong(seed), mask.contiguous().view(torch.uint8),
                  picks, device=coords.device)
        return gather(masked - mean.view(bsz, -1, 1), picks), mean, mask
    for i in range(bsz):
        cand = mask[i].nonzero().view(-1)
        nc = cand.numel()
        if nc >= k:
            sel = np.random.choice(nc, k, replace=False)
        elif nc > 0:
            sel = np.concatenate([np.arange(nc).repeat(k // nc), np.random.choice(nc, k % nc, replace=False)])
            np.random.shuffle(sel)
        else:
            continue
        picks[i] = cand[sel]
    return gather(masked - mean.view(bsz, -1, 1), picks), mean, mask


class _NeighborInterpolate(Function):
    """modules/functional/interpolatation.py:8-35: gradient only w.r.t. centers_features."""

    @staticmethod
    def forward(ctx, points_coords, centers_coords, centers_features):
        centers_coords = centers_coords.contiguous()
        points_coords = points_coords.contiguous()
        centers_features = centers_features.contiguous()
        out, indices, weights = _backend.three_nearest_neighbors_interpolate_forward(
            points_coords, centers_coords, centers_features)
        ctx.save_for_backward(indices, weights)
        ctx.num_centers = centers_coords.size(-1)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        indices, weights = ctx.saved_tensors
        g = _backend.three_nearest_neighbors_interpolate_backward(grad_output.contiguous(), indices, weights,
                                                                  ctx.num_centers)
        return None, None, g


nearest_neighbor_interpolate = _NeighborInterpolate.apply


def vox_stats_mode():
    """PVCNN_B200_VOX = exact (default) | aten | fused.
    exact: per-cloud mean by the reference's own ATen call (coords.mean(2), modules/voxelization.py:18), the
           max-norm and the element-wise tail in our kernels (bit-identical to the reference's op sequence);
    aten:  mean AND denominator by the reference's ATen calls (:18-20), element-wise tail in our kernel;
    fused: single kernel with an fp64 mean (differs from torch's reduction by an ulp on rare .5 ties)."""
    m = os.environ.get("PVCNN_B200_VOX", "exact").lower()
    return m if m in ("exact", "aten", "fused") else "exact"


def voxel_stats(coords, normalize, eps, mode=None):
    """(vox_stats flag, mean [B,3], denom [B] or None) for pvcnn_voxelize_* / the fused block."""
    mode = mode or vox_stats_mode()
    if mode == "fused":
        return 0, None, None
    mean = coords.mean(2, keepdim=True)                                     # modules/voxelization.py:18
    if not normalize:
        return 1, mean.reshape(-1, 3).contiguous(), None
    if mode == "aten":
        nc = coords - mean
        denom = nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps   # :20
        return 2, mean.reshape(-1, 3).contiguous(), denom.reshape(-1).contiguous()
    denom = torch.empty(coords.shape[0], dtype=torch.float32, device=coords.device)
    return 1, mean.reshape(-1, 3).contiguous(), denom


def voxelize_coords(coords, resolution, normalize=True, eps=0.0):
    """Replacement for the tensor ops of modules/voxelization.py:17-24: returns
    (norm_coords float [B,3,N] clamped to [0,r-1], vox_coords int32 [B,3,N]); bit-identical to the reference's
    op sequence on the same device (see vox_stats_mode)."""
    coords = coords.detach().contiguous().float()
    b, _, n = coords.shape
    nc = torch.empty_like(coords)
    vc = torch.empty(coords.shape, dtype=torch.int32, device=coords.device)
    flag, mean, denom = voxel_stats(coords, normalize, eps)
    if flag == 0:
        _lib.call("pvcnn_voxelize_coords", b, n, int(resolution), bool(normalize), float(eps), coords, nc, vc)
        return nc, vc
    if normalize and flag == 1:
        _lib.call("pvcnn_voxelize_denom", b, n, float(eps), coords, mean, denom)
    _lib.call("pvcnn_voxelize_apply", b, n, int(resolution), bool(normalize), coords, mean, denom, nc, vc)
    return nc, vc
