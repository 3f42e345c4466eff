"""`modules.functional` alias (reference: modules/functional/__init__.py:1-7)."""
from pvcnn_b200.functional import (avg_voxelize, trilinear_devoxelize, ball_query, grouping, gather,
                                   furthest_point_sample, logits_mask, nearest_neighbor_interpolate, kl_loss,
                                   huber_loss)

__all__ = ["avg_voxelize", "trilinear_devoxelize", "ball_query", "grouping", "gather", "furthest_point_sample",
           "logits_mask", "nearest_neighbor_interpolate", "kl_loss", "huber_loss"]
