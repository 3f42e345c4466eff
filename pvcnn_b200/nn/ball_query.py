import torch.nn as nn

from .. import functional as F


class BallQuery(nn.Module):
    """Radius neighbourhood grouper (reference: modules/ball_query.py:9-34).
    forward(points_coords [B,3,N], centers_coords [B,3,M], points_features [B,C,N] | None)
    -> [B, (3+)C, M, U]: neighbour coordinates relative to their centre, optionally stacked on
    the neighbours' features."""

    def __init__(self, radius, num_neighbors, include_coordinates=True):
        super().__init__()
        self.radius = radius
        self.num_neighbors = num_neighbors
        self.include_coordinates = include_coordinates

    def forward(self, points_coords, centers_coords, points_features=None):
        points_coords = points_coords.contiguous()
        centers_coords = centers_coords.contiguous()
        idx = F.ball_query(centers_coords, points_coords, self.radius, self.num_neighbors)
        if points_features is None:
            assert self.include_coordinates, "No Features For Grouping"
            return F.group_concat(points_coords, centers_coords, None, idx)
        if not self.include_coordinates:
            return F.grouping(points_features, idx)
        # one kernel writes [B,3+C,M,U] directly (the reference materialises three intermediates and a cat)
        return F.group_concat(points_coords, centers_coords, points_features, idx)

    def extra_repr(self):
        return "radius={}, num_neighbors={}{}".format(
            self.radius, self.num_neighbors, ", include coordinates" if self.include_coordinates else "")
