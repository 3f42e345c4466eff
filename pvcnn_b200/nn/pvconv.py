import os

import torch.nn as nn

from .. import functional as F
from .se import SE3d
from .shared_mlp import SharedMLP
from .voxelization import Voxelization


class PVConv(nn.Module):
    """Point-Voxel convolution block (reference: modules/pvconv.py:11-39).

        voxel branch : avg_voxelize -> (Conv3d, BN3d(eps=1e-4), LeakyReLU(0.1)) x2 [-> SE3d] -> trilinear_devoxelize
        point branch : SharedMLP (Conv1d k=1, BN1d, ReLU)
        output       : voxel branch + point branch,  forward((features, coords)) -> (fused, coords)

    The parameters live in the same sub-modules, under the same names and shapes, as in the
    reference (`voxel_layers.{0,1,3,4[,6]}`, `point_features.layers.{0,1}`), so
    reference checkpoints load unchanged.  Execution, however, does not go through those
    sub-modules' forward(): the whole block runs as one fused sm_100a pipeline
    (pvcnn_b200/fused.py -> libpvcnn_b200.so) with channels-last grids, tcgen05 implicit-GEMM
    convolutions and BatchNorm/activation folded into the neighbouring kernels.

    PVCNN_B200_PVCONV=composed selects a bring-up path that chains the stand-alone ops
    (F.avg_voxelize / F.trilinear_devoxelize) around torch's dense layers; it exists for
    debugging and as the in-repo comparison arm, not as a product path.
    """

    def __init__(self, in_channels, out_channels, kernel_size, resolution, with_se=False, normalize=True, eps=0):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.resolution = resolution
        self.with_se = with_se

        self.voxelization = Voxelization(resolution, normalize=normalize, eps=eps)
        pad = kernel_size // 2
        stack = []
        cin = in_channels
        for _ in range(2):
            stack += [nn.Conv3d(cin, out_channels, kernel_size, stride=1, padding=pad),
                      nn.BatchNorm3d(out_channels, eps=1e-4), nn.LeakyReLU(0.1, True)]
            cin = out_channels
        if with_se:
            stack.append(SE3d(out_channels))
        self.voxel_layers = nn.Sequential(*stack)
        self.point_features = SharedMLP(in_channels, out_channels)

    def _forward_composed(self, features, coords):
        grid, norm_coords = self.voxelization(features, coords)
        grid = self.voxel_layers(grid)
        vox = F.trilinear_devoxelize(grid, norm_coords, self.resolution, self.training)
        return vox + self.point_features(features), coords

    def _fused_unsupported_reason(self):
        """The fused pipeline hard-wires what the reference's PVConv always builds (modules/pvconv.py:19-31): 3x3x3 convs
        with padding 1, affine BatchNorm with running statistics, ONE LeakyReLU slope, one BN eps per branch.  Anything
        else (kernel_size 1/5, a module someone edited after construction) must not reach kernels that would read the
        weights with the wrong layout."""
        c1, n1, a1, c2, n2, a2 = (self.voxel_layers[i] for i in range(6))
        if self.kernel_size != 3:
            return "kernel_size=%r (the tensor-core pipeline implements 3x3x3 / padding 1)" % (self.kernel_size,)
        for conv, cin in ((c1, self.in_channels), (c2, self.out_channels)):
            if tuple(conv.weight.shape) != (self.out_channels, cin, 3, 3, 3) or conv.bias is None:
                return "unexpected Conv3d weight shape %s / missing bias" % (tuple(conv.weight.shape),)
            if tuple(conv.stride) != (1, 1, 1) or tuple(conv.padding) != (1, 1, 1) or conv.groups != 1:
                return "Conv3d stride/padding/groups differ from (1, 1, 1)"
        pt = self.point_features.layers
        if len(pt) != 3 or tuple(pt[0].weight.shape) != (self.out_channels, self.in_channels, 1) or pt[0].bias is None:
            return "point branch is not a single (Conv1d k=1, BatchNorm1d, ReLU) layer"
        for bn in (n1, n2, pt[1]):
            if not bn.affine or not bn.track_running_stats or bn.running_mean is None:
                return "BatchNorm without affine parameters / running statistics"
        if n1.eps != n2.eps or a1.negative_slope != a2.negative_slope:
            return "the two voxel layers use different BatchNorm eps / LeakyReLU slopes"
        if self.with_se and self.out_channels < 8:
            return "SE3d with fewer than 8 channels"
        return None

    def forward(self, inputs):
        features, coords = inputs
        mode = os.environ.get("PVCNN_B200_PVCONV", "fused")
        if mode == "composed":
            return self._forward_composed(features, coords)
        why = self._fused_unsupported_reason()
        if why is not None:
            if mode == "fused_strict":
                raise RuntimeError("PVConv: configuration outside the fused sm_100a pipeline: " + why)
            return self._forward_composed(features, coords)   # stand-alone sm_100a ops + torch dense layers
        from ..fused import pvconv_fused
        return pvconv_fused(self, features, coords), coords
