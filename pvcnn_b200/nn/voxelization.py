import torch.nn as nn

from .. import functional as F


class Voxelization(nn.Module):
    """Point cloud -> R^3 average-pooled grid (reference: modules/voxelization.py:9-28).

    forward(features [B,C,N], coords [B,3,N]) -> (grid [B,C,R,R,R], norm_coords [B,3,N]).
    The ~9 ATen kernels of the reference's coordinate normalisation are one fused kernel
    (pvcnn_voxelize_coords); the scatter-mean is pvcnn_avg_voxelize."""

    def __init__(self, resolution, normalize=True, eps=0):
        super().__init__()
        self.r = int(resolution)
        self.normalize = normalize
        self.eps = eps

    def forward(self, features, coords):
        norm_coords, vox_coords = F.voxelize_coords(coords, self.r, self.normalize, self.eps)
        return F.avg_voxelize(features, vox_coords, self.r), norm_coords

    def extra_repr(self):
        s = "resolution={}".format(self.r)
        if self.normalize:
            s += ", normalized eps = {}".format(self.eps)
        return s
