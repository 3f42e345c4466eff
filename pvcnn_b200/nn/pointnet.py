import torch
import torch.nn as nn

from .. import functional as F
from .ball_query import BallQuery
from .shared_mlp import SharedMLP


def _as_nested(widths, count):
    """Normalise `out_channels` to one list of widths per branch (modules/pointnet.py:14-17, :56-60)."""
    if not isinstance(widths, (list, tuple)):
        return [[widths]] * count
    if not isinstance(widths[0], (list, tuple)):
        return [widths] * count
    return widths


class PointNetAModule(nn.Module):
    """Global ("all points") abstraction (reference: modules/pointnet.py:11-46): per-branch
    SharedMLP followed by a max over points; returns a single zero centre."""

    def __init__(self, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        branches = _as_nested(out_channels, 1)
        cin = in_channels + (3 if include_coordinates else 0)
        self.mlps = nn.ModuleList([SharedMLP(in_channels=cin, out_channels=w, dim=1) for w in branches])
        self.include_coordinates = include_coordinates
        self.out_channels = sum(w[-1] for w in branches)

    def forward(self, inputs):
        features, coords = inputs
        origin = torch.zeros((coords.size(0), 3, 1), device=coords.device)
        from .shared_mlp import native_mlp_enabled
        if features.is_cuda and native_mlp_enabled():
            from .. import mlp as _mlp
            if all(_mlp.native_supported(m.layers) for m in self.mlps):
                # features (+ coordinates) go side by side into channels-last rows ONCE, the tensor-core MLP runs on them,
                # and the max over ALL points is folded into the last layer's BatchNorm pass (segmented over the chip):
                # neither the concat nor the [B, C', N] activation is written
                b, _, n = features.shape
                rows, lo = _mlp.cat_cl([features, coords] if self.include_coordinates else [features], n)
                pooled = [_mlp._run(_mlp._FromCL, _mlp.mlp_cl(m.layers, rows, lo, pool_u=n), b,
                                             list(m.layers)[-3].out_channels, 1) for m in self.mlps]
                return (torch.cat(pooled, dim=1) if len(pooled) > 1 else pooled[0]), origin
        if self.include_coordinates:
            features = torch.cat([features, coords], dim=1)
        pooled = [mlp(features).max(dim=-1, keepdim=True).values for mlp in self.mlps]
        return (torch.cat(pooled, dim=1) if len(pooled) > 1 else pooled[0]), origin

    def extra_repr(self):
        return f"out_channels={self.out_channels}, include_coordinates={self.include_coordinates}"


class PointNetSAModule(nn.Module):
    """Set abstraction (reference: modules/pointnet.py:49-92): FPS centres, one
    (BallQuery -> SharedMLP(dim=2) -> max over neighbours) branch per radius."""

    def __init__(self, num_centers, radius, num_neighbors, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        radii = list(radius) if isinstance(radius, (list, tuple)) else [radius]
        ks = list(num_neighbors) if isinstance(num_neighbors, (list, tuple)) else [num_neighbors] * len(radii)
        assert len(radii) == len(ks)
        branches = _as_nested(out_channels, len(radii))
        assert len(radii) == len(branches)
        cin = in_channels + (3 if include_coordinates else 0)
        self.groupers = nn.ModuleList(
            [BallQuery(radius=r, num_neighbors=k, include_coordinates=include_coordinates) for r, k in zip(radii, ks)])
        self.mlps = nn.ModuleList([SharedMLP(in_channels=cin, out_channels=w, dim=2) for w in branches])
        self.num_centers = num_centers
        self.out_channels = sum(w[-1] for w in branches)

    def _branch(self, grouper, mlp, coords, centers, features):
        """One (BallQuery -> SharedMLP(dim=2) -> max over neighbours) branch.  On CUDA the grouping writes channels-last
        rows that feed the tensor-core MLP directly and the max over U is folded into the last layer's BatchNorm pass:
        neither the [B,C+3,M,U] input nor the [B,C',M,U] output of the reference is materialised."""
        from .shared_mlp import native_mlp_enabled
        if coords.is_cuda and native_mlp_enabled() and grouper.include_coordinates:
            from .. import mlp as _mlp
            if _mlp.native_supported(mlp.layers):
                idx = F.ball_query(centers.contiguous(), coords.contiguous(), grouper.radius, grouper.num_neighbors)
                return _mlp.sa_branch(mlp.layers, coords, centers, features, idx)
        return mlp(grouper(coords, centers, features)).max(dim=-1).values

    def forward(self, inputs):
        features, coords = inputs
        centers = F.furthest_point_sample(coords, self.num_centers)
        pooled = [self._branch(g, mlp, coords, centers, features) for g, mlp in zip(self.groupers, self.mlps)]
        return (torch.cat(pooled, dim=1) if len(pooled) > 1 else pooled[0]), centers

    def extra_repr(self):
        return f"num_centers={self.num_centers}, out_channels={self.out_channels}"


class PointNetFPModule(nn.Module):
    """Feature propagation (reference: modules/pointnet.py:95-111): 3-NN inverse-distance
    interpolation of the centres' features onto the points, optional skip concat, SharedMLP."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.mlp = SharedMLP(in_channels=in_channels, out_channels=out_channels, dim=1)

    def forward(self, inputs):
        points_coords, centers_coords, centers_features = inputs[:3]
        skip = inputs[3] if len(inputs) > 3 else None
        x = F.nearest_neighbor_interpolate(points_coords, centers_coords, centers_features)
        from .shared_mlp import native_mlp_enabled
        if x.is_cuda and native_mlp_enabled():
            from .. import mlp as _mlp
            if _mlp.native_supported(self.mlp.layers):
                # interpolated features and the skip connection are written side by side into channels-last rows that
                # feed the tensor-core MLP directly (no [B, C1+C2, N] concat, no separate layout conversion)
                b, _, n = x.shape
                rows, lo = _mlp.cat_cl([x] if skip is None else [x, skip], n)
                z = _mlp.mlp_cl(self.mlp.layers, rows, lo)
                return _mlp._run(_mlp._FromCL, z, b, list(self.mlp.layers)[-3].out_channels, n), points_coords
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.mlp(x), points_coords
