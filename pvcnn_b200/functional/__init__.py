"""`modules.functional` surface of the reference (modules/functional/__init__.py:1-7), served by
the sm_100a library.  Every public name keeps the reference's positional signature, dtype
coercions and autograd behaviour (file:line cited per symbol)."""
from .ops import (avg_voxelize, trilinear_devoxelize, ball_query, grouping, gather, furthest_point_sample,
                  logits_mask, nearest_neighbor_interpolate, voxelize_coords, group_concat)
from .loss import kl_loss, huber_loss

__all__ = ["avg_voxelize", "trilinear_devoxelize", "ball_query", "grouping", "gather", "furthest_point_sample",
           "logits_mask", "nearest_neighbor_interpolate", "kl_loss", "huber_loss", "voxelize_coords", "group_concat"]
