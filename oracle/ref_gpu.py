"""GPU reference arm (TEST / BENCH INFRASTRUCTURE ONLY): the reference's PVConv block executed
with the reference's OWN CUDA kernels (oracle/_ref/_pvcnn_backend.so, built unmodified from
/root/reference by oracle/build_ref.py) and torch's cuDNN/cuBLAS dense layers -- i.e. exactly
what `modules.PVConv` of mit-han-lab/pvcnn runs (modules/pvconv.py:33-39), re-wired here because the
reference's Python package cannot travel to the GPU box.  Never imported by the product path."""
import importlib.util
import os

import torch
import torch.nn as nn
from torch.autograd import Function

_HERE = os.path.dirname(os.path.abspath(__file__))
_BACKEND = None


def backend():
    global _BACKEND
    if _BACKEND is None:
        path = os.path.join(_HERE, "_ref", "_pvcnn_backend.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        spec = importlib.util.spec_from_file_location("_pvcnn_backend", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _BACKEND = mod
    return _BACKEND


class _RefVox(Function):  # modules/functional/voxelization.py:8-37
    @staticmethod
    def forward(ctx, f, c, r):
        out, ind, cnt = backend().avg_voxelize_forward(f.contiguous(), c.int().contiguous(), r)
        ctx.save_for_backward(ind, cnt)
        return out.view(f.shape[0], f.shape[1], r, r, r)

    @staticmethod
    def backward(ctx, g):
        ind, cnt = ctx.saved_tensors
        b, c = g.shape[:2]
        return backend().avg_voxelize_backward(g.contiguous().view(b, c, -1), ind, cnt), None, None


class _RefDevox(Function):  # modules/functional/devoxelization.py:8-39
    @staticmethod
    def forward(ctx, f, c, r, training):
        b, ch = f.shape[:2]
        outs, inds, wgts = backend().trilinear_devoxelize_forward(r, training, c.contiguous(),
                                                                 f.contiguous().view(b, ch, -1))
        if training:
            ctx.save_for_backward(inds, wgts)
            ctx.r = r
        return outs

    @staticmethod
    def backward(ctx, g):
        inds, wgts = ctx.saved_tensors
        gi = backend().trilinear_devoxelize_backward(g.contiguous(), inds, wgts, ctx.r)
        return gi.view(g.size(0), g.size(1), ctx.r, ctx.r, ctx.r), None, None, None


class RefPVConv(nn.Module):
    """Same sub-module names as the reference, so state_dicts are interchangeable."""

    def __init__(self, in_channels, out_channels, kernel_size, resolution, with_se=False, normalize=True, eps=0):
        super().__init__()
        self.r, self.normalize, self.eps = int(resolution), normalize, eps
        layers = [nn.Conv3d(in_channels, out_channels, kernel_size, stride=1, padding=kernel_size // 2),
                  nn.BatchNorm3d(out_channels, eps=1e-4), nn.LeakyReLU(0.1, True),
                  nn.Conv3d(out_channels, out_channels, kernel_size, stride=1, padding=kernel_size // 2),
                  nn.BatchNorm3d(out_channels, eps=1e-4), nn.LeakyReLU(0.1, True)]
        if with_se:
            from pvcnn_b200.nn.se import SE3d  # same tiny torch module; parameters only
            layers.append(SE3d(out_channels))
        self.voxel_layers = nn.Sequential(*layers)
        pf = nn.Module()
        pf.layers = nn.Sequential(nn.Conv1d(in_channels, out_channels, 1), nn.BatchNorm1d(out_channels), nn.ReLU(True))
        self.point_features = pf

    def forward(self, inputs):
        features, coords = inputs
        c = coords.detach()
        nc = c - c.mean(2, keepdim=True)  # modules/voxelization.py:17-24
        if self.normalize:
            nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + self.eps) + 0.5
        else:
            nc = (nc + 1) / 2.0
        nc = torch.clamp(nc * self.r, 0, self.r - 1)
        vc = torch.round(nc).to(torch.int32)
        grid = _RefVox.apply(features, vc, self.r)
        grid = self.voxel_layers(grid)
        vox = _RefDevox.apply(grid, nc, self.r, self.training)
        return vox + self.point_features.layers(features), coords
