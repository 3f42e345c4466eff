"""Tensor-level backend: the 12 functions the reference exposes as `_pvcnn_backend`
(modules/functional/src/bindings.cpp:10-37), re-implemented as thin allocators around the C ABI
(include/pvcnn_b200.h).  Checks, shapes and return conventions follow the reference's C++
wrappers (file:line cited per function); outputs are allocated with torch.empty because the
library initialises them itself.
"""
import torch

from .. import _lib
from ._checks import check_f, check_i


def avg_voxelize_forward(features, coords, resolution):
    """voxelization/vox.cpp:17-43 -> (out [B,C,R^3], ind [B,N], cnt [B,R^3])"""
    check_f(features, "features"); check_i(coords, "coords")
    b, c, n = features.shape
    r = int(resolution)
    r2, r3 = r * r, r * r * r
    dev = features.device
    ind = torch.empty((b, n), dtype=torch.int32, device=dev)
    out = torch.empty((b, c, r3), dtype=torch.float32, device=dev)
    cnt = torch.empty((b, r3), dtype=torch.int32, device=dev)
    _lib.call("pvcnn_avg_voxelize", b, c, n, r, r2, r3, coords, features, ind, cnt, out)
    return out, ind, cnt


def avg_voxelize_backward(grad_y, indices, cnt):
    """voxelization/vox.cpp:54-76 -> grad_x [B,C,N]"""
    check_f(grad_y, "grad_y"); check_i(indices, "indices"); check_i(cnt, "cnt")
    b, c, s = grad_y.shape
    n = indices.shape[1]
    grad_x = torch.empty((b, c, n), dtype=torch.float32, device=grad_y.device)
    _lib.call("pvcnn_avg_voxelize_grad", b, c, n, s, indices, cnt, grad_y, grad_x)
    return grad_x


def trilinear_devoxelize_forward(r, is_training, coords, features):
    """interpolate/trilinear_devox.cpp:18-55 -> (outs [B,C,N], inds [B,8,N], wgts [B,8,N]);
    inds/wgts are dummy [1] tensors when not training (:45-53)."""
    check_f(features, "features"); check_f(coords, "coords")
    b, c = features.shape[:2]
    n = coords.shape[2]
    r = int(r)
    dev = features.device
    outs = torch.empty((b, c, n), dtype=torch.float32, device=dev)
    if is_training:
        inds = torch.empty((b, 8, n), dtype=torch.int32, device=dev)
        wgts = torch.empty((b, 8, n), dtype=torch.float32, device=dev)
        _lib.call("pvcnn_trilinear_devoxelize", b, c, n, r, r * r, r * r * r, 1, coords, features, inds, wgts, outs)
    else:
        inds = torch.zeros((1,), dtype=torch.int32, device=dev)
        wgts = torch.zeros((1,), dtype=torch.float32, device=dev)
        _lib.call("pvcnn_trilinear_devoxelize", b, c, n, r, r * r, r * r * r, 0, coords, features, None, None, outs)
    return outs, inds, wgts


def trilinear_devoxelize_backward(grad_y, indices, weights, r):
    """interpolate/trilinear_devox.cpp:67-91 -> grad_x [B,C,R^3]"""
    check_f(grad_y, "grad_y"); check_f(weights, "weights"); check_i(indices, "indices")
    b, c, n = grad_y.shape
    r3 = int(r) ** 3
    grad_x = torch.empty((b, c, r3), dtype=torch.float32, device=grad_y.device)
    _lib.call("pvcnn_trilinear_devoxelize_grad", b, c, n, r3, indices, weights, grad_y, grad_x)
    return grad_x


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    """ball_query/ball_query.cpp:6-30 -> int [B,M,U]; the kernel receives radius*radius in fp32 (:24)."""
    check_f(centers_coords, "centers_coords"); check_f(points_coords, "points_coords")
    b, _, m = centers_coords.shape
    n = points_coords.shape[2]
    u = int(num_neighbors)
    out = torch.empty((b, m, u), dtype=torch.int32, device=points_coords.device)
    r2 = float(torch.tensor(float(radius), dtype=torch.float32) * torch.tensor(float(radius), dtype=torch.float32))
    _lib.call("pvcnn_ball_query", b, n, m, r2, u, centers_coords, points_coords, out)
    return out


def grouping_forward(features, indices):
    """grouping/grouping.cpp:6-24 -> [B,C,M,U]"""
    check_f(features, "features"); check_i(indices, "indices")
    b, c, n = features.shape
    _, m, u = indices.shape
    out = torch.empty((b, c, m, u), dtype=torch.float32, device=features.device)
    _lib.call("pvcnn_grouping", b, c, n, m, u, features, indices, out)
    return out


def grouping_backward(grad_y, indices, n):
    """grouping/grouping.cpp:26-44 -> [B,C,N]"""
    check_f(grad_y, "grad_y"); check_i(indices, "indices")
    b, c = grad_y.shape[:2]
    _, m, u = indices.shape
    grad_x = torch.empty((b, c, int(n)), dtype=torch.float32, device=grad_y.device)
    _lib.call("pvcnn_grouping_grad", b, c, int(n), m, u, grad_y, indices, grad_x)
    return grad_x


def group_concat_forward(points_coords, centers_coords, features, indices):
    """Fused modules/ball_query.py:16-30 -> [B,3+C,M,U] (features may be None: C = 0)."""
    check_f(points_coords, "points_coords"); check_f(centers_coords, "centers_coords"); check_i(indices, "indices")
    if features is not None:
        check_f(features, "features")
    b, _, n = points_coords.shape
    _, m, u = indices.shape
    c = 0 if features is None else features.shape[1]
    out = torch.empty((b, 3 + c, m, u), dtype=torch.float32, device=points_coords.device)
    _lib.call("pvcnn_group_concat", b, c, n, m, u, points_coords, centers_coords, features, indices, out)
    return out


def group_concat_backward(grad_y, indices, n, need_features=True, need_points=False, need_centers=False):
    """-> (grad_features [B,C,N] | None, grad_points_coords [B,3,N] | None, grad_centers_coords [B,3,M] | None)"""
    check_f(grad_y, "grad_y"); check_i(indices, "indices")
    b, ct = grad_y.shape[:2]
    _, m, u = indices.shape
    c = ct - 3
    dev = grad_y.device
    gf = torch.empty((b, c, int(n)), dtype=torch.float32, device=dev) if c > 0 else None
    gp = torch.empty((b, 3, int(n)), dtype=torch.float32, device=dev) if need_points else None
    gc = torch.empty((b, 3, m), dtype=torch.float32, device=dev) if need_centers else None
    _lib.call("pvcnn_group_concat_grad", b, c, int(n), m, u, grad_y, indices, gf, gp, gc)
    return (gf if need_features else None), gp, gc


def gather_features_forward(features, indices):
    """sampling/sampling.cpp:6-23 -> [B,C,M]"""
    check_f(features, "features"); check_i(indices, "indices")
    b, c, n = features.shape
    m = indices.shape[1]
    out = torch.empty((b, c, m), dtype=torch.float32, device=features.device)
    _lib.call("pvcnn_gather_features", b, c, n, m, features, indices, out)
    return out


def gather_features_backward(grad_y, indices, n):
    """sampling/sampling.cpp:25-41 -> [B,C,N]"""
    check_f(grad_y, "grad_y"); check_i(indices, "indices")
    b, c = grad_y.shape[:2]
    m = indices.shape[1]
    grad_x = torch.empty((b, c, int(n)), dtype=torch.float32, device=grad_y.device)
    _lib.call("pvcnn_gather_features_grad", b, c, int(n), m, grad_y, indices, grad_x)
    return grad_x


def furthest_point_sampling(coords, num_samples):
    """sampling/sampling.cpp:43-58 -> int [B,M]"""
    check_f(coords, "coords")
    b, _, n = coords.shape
    m = int(num_samples)
    indices = torch.zeros((b, m), dtype=torch.int32, device=coords.device)
    _lib.call("pvcnn_furthest_point_sampling", b, n, m, coords, None, indices)
    return indices


def three_nearest_neighbors_interpolate_forward(points_coords, centers_coords, centers_features):
    """interpolate/neighbor_interpolate.cpp:6-40 -> (out [B,C,N], indices [B,3,N], weights [B,3,N])"""
    check_f(points_coords, "points_coords"); check_f(centers_coords, "centers_coords")
    check_f(centers_features, "centers_features")
    b, c, m = centers_features.shape
    n = points_coords.shape[2]
    dev = points_coords.device
    indices = torch.empty((b, 3, n), dtype=torch.int32, device=dev)
    weights = torch.empty((b, 3, n), dtype=torch.float32, device=dev)
    out = torch.empty((b, c, n), dtype=torch.float32, device=dev)
    _lib.call("pvcnn_three_nearest_neighbors_interpolate", b, c, m, n, points_coords, centers_coords,
              centers_features, indices, weights, out)
    return out, indices, weights


def three_nearest_neighbors_interpolate_backward(grad_y, indices, weights, m):
    """interpolate/neighbor_interpolate.cpp:42-65 -> [B,C,M]"""
    check_f(grad_y, "grad_y"); check_i(indices, "indices"); check_f(weights, "weights")
    b, c, n = grad_y.shape
    grad_x = torch.empty((b, c, int(m)), dtype=torch.float32, device=grad_y.device)
    _lib.call("pvcnn_three_nearest_neighbors_interpolate_grad", b, c, n, int(m), grad_y, indices, weights, grad_x)
    return grad_x
