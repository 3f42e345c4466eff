"""TEST INFRASTRUCTURE (checker only; never imported by pvcnn_b200/ or modules/).

CPU restatement of the reference's PointNet++ style modules for parity tests of the native paths:
  SharedMLP          modules/shared_mlp.py:6-33
  BallQuery          modules/ball_query.py:9-34
  PointNetAModule    modules/pointnet.py:11-46
  PointNetSAModule   modules/pointnet.py:49-92
  PointNetFPModule   modules/pointnet.py:95-111
The index decisions (FPS picks, ball-query neighbour lists, 3-NN indices and fp32 inverse-distance weights) come from
the C oracle (oracle/pvcnn_oracle.c, pinned against the reference's kernels by tests/golden); the dense arithmetic is the
same torch calls the reference makes (nn.Conv1d/Conv2d, nn.BatchNorm, nn.ReLU, max, cat), evaluated in the module's dtype
(float64 = truth).  State-dict names equal the reference's, so a product module's state_dict loads directly.
"""
import numpy as np
import torch
import torch.nn as nn

from . import oracle as O


def _np32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)


def _group(features, idx_t):
    """features [B,C,N], idx [B,M,U] (long) -> [B,C,M,U]   (grouping.cu:33, pure index copy; differentiable)"""
    b, c, n = features.shape
    _, m, u = idx_t.shape
    flat = idx_t.reshape(b, 1, m * u).expand(-1, c, -1)
    return torch.gather(features, 2, flat).view(b, c, m, u)


class SharedMLP(nn.Module):
    def __init__(self, in_channels, out_channels, dim=1):
        super().__init__()
        conv, bn = (nn.Conv1d, nn.BatchNorm1d) if dim == 1 else (nn.Conv2d, nn.BatchNorm2d)
        if not isinstance(out_channels, (list, tuple)):
            out_channels = [out_channels]
        layers = []
        for oc in out_channels:
            layers += [conv(in_channels, oc, 1), bn(oc), nn.ReLU(True)]
            in_channels = oc
        self.layers = nn.Sequential(*layers)

    def forward(self, inputs):
        if isinstance(inputs, (list, tuple)):
            return (self.layers(inputs[0]), *inputs[1:])
        return self.layers(inputs)


class BallQuery(nn.Module):
    def __init__(self, radius, num_neighbors, include_coordinates=True):
        super().__init__()
        self.radius, self.num_neighbors, self.include_coordinates = radius, num_neighbors, include_coordinates

    def forward(self, points_coords, centers_coords, points_features=None):
        idx = O.ball_query(_np32(centers_coords), _np32(points_coords), self.radius, self.num_neighbors)
        idx_t = torch.from_numpy(idx).long()
        nc = _group(points_coords, idx_t) - centers_coords.unsqueeze(-1)
        if points_features is None:
            return nc
        nf = _group(points_features, idx_t)
        return torch.cat([nc, nf], dim=1) if self.include_coordinates else nf


def furthest_point_sample(coords, m):
    idx = torch.from_numpy(O.furthest_point_sampling(_np32(coords), m)).long()
    return torch.gather(coords, 2, idx.unsqueeze(1).expand(-1, 3, -1))


def nearest_neighbor_interpolate(points_coords, centers_coords, centers_features):
    idx, w = O.three_nn(_np32(points_coords), _np32(centers_coords))     # fp32 decisions / weights, as the reference
    idx_t = torch.from_numpy(idx).long()
    w_t = torch.from_numpy(w).to(centers_features.dtype)
    b, c, m = centers_features.shape
    n = idx_t.shape[2]
    out = 0
    for k in range(3):
        out = out + torch.gather(centers_features, 2, idx_t[:, k:k + 1, :].expand(-1, c, -1)) * w_t[:, k:k + 1, :]
    return out


def _nested(out_channels, count):
    if not isinstance(out_channels, (list, tuple)):
        return [[out_channels]] * count
    if not isinstance(out_channels[0], (list, tuple)):
        return [out_channels] * count
    return out_channels


class PointNetAModule(nn.Module):
    def __init__(self, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        branches = _nested(out_channels, 1)
        self.mlps = nn.ModuleList([SharedMLP(in_channels + (3 if include_coordinates else 0), w, dim=1) for w in branches])
        self.include_coordinates = include_coordinates

    def forward(self, inputs):
        features, coords = inputs
        if self.include_coordinates:
            features = torch.cat([features, coords], dim=1)
        origin = torch.zeros((coords.size(0), 3, 1), dtype=coords.dtype)
        outs = [mlp(features).max(dim=-1, keepdim=True).values for mlp in self.mlps]
        return (torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]), origin


class PointNetSAModule(nn.Module):
    def __init__(self, num_centers, radius, num_neighbors, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        radius = list(radius) if isinstance(radius, (list, tuple)) else [radius]
        if not isinstance(num_neighbors, (list, tuple)):
            num_neighbors = [num_neighbors] * len(radius)
        branches = _nested(out_channels, len(radius))
        self.groupers = nn.ModuleList([BallQuery(r, k, include_coordinates) for r, k in zip(radius, num_neighbors)])
        self.mlps = nn.ModuleList([SharedMLP(in_channels + (3 if include_coordinates else 0), w, dim=2) for w in branches])
        self.num_centers = num_centers

    def forward(self, inputs):
        features, coords = inputs
        centers = furthest_point_sample(coords, self.num_centers)
        outs = [mlp(g(coords, centers, features)).max(dim=-1).values for g, mlp in zip(self.groupers, self.mlps)]
        return (torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]), centers


class PointNetFPModule(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.mlp = SharedMLP(in_channels, out_channels, dim=1)

    def forward(self, inputs):
        points_coords, centers_coords, centers_features = inputs[:3]
        skip = inputs[3] if len(inputs) > 3 else None
        x = nearest_neighbor_interpolate(points_coords, centers_coords, centers_features)
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.mlp(x), points_coords


def clone_as_oracle(product_module, oracle_module, dtype=torch.float64):
    """Load the product module's parameters / buffers into the oracle module (same state_dict names) in `dtype`."""
    oracle_module = oracle_module.to(dtype)
    sd = {k: v.detach().cpu().to(dtype) if v.is_floating_point() else v.detach().cpu()
          for k, v in product_module.state_dict().items()}
    oracle_module.load_state_dict(sd)
    oracle_module.train(product_module.training)
    return oracle_module
