"""Frustum-PointNet loss and box-corner helper (reference: modules/frustum.py:11-124).
Pure torch, outside the accelerated hot path; restated so that
`modules.frustum` / `modules.FrustumPointNetLoss` stay importable for the reference's
configs (configs/kitti/frustum/__init__.py:8) and meters (meters/kitti/frustum.py:4)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as TF

from ..functional.loss import huber_loss

# corner sign pattern, counter-clockwise, top face first (modules/frustum.py:103-105)
_SX = (1, 1, -1, -1, 1, 1, -1, -1)
_SY = (1, 1, 1, 1, -1, -1, -1, -1)
_SZ = (1, -1, -1, 1, 1, -1, -1, 1)


def get_box_corners_3d(centers, headings, sizes, with_flip=False):
    """centers [N,3], headings [N], sizes [N,3]=(l,w,h) -> corners [N,3,8] (and the heading+pi
    flipped box when with_flip)."""
    half = sizes / 2
    l, w, h = half[:, 0], half[:, 1], half[:, 2]
    sx = centers.new_tensor(_SX); sy = centers.new_tensor(_SY); sz = centers.new_tensor(_SZ)
    local = torch.stack([l[:, None] * sx, h[:, None] * sy, w[:, None] * sz], dim=1)  # [N,3,8]
    c, s = torch.cos(headings), torch.sin(headings)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    rot = torch.stack([c, zero, s, zero, one, zero, -s, zero, c], dim=1).view(-1, 3, 3)
    shift = centers.unsqueeze(-1)
    if not with_flip:
        return rot @ local + shift
    rot_flip = torch.stack([-c, zero, -s, zero, one, zero, s, zero, -c], dim=1).view(-1, 3, 3)
    return rot @ local + shift, rot_flip @ local + shift


class FrustumPointNetLoss(nn.Module):
    def __init__(self, num_heading_angle_bins, num_size_templates, size_templates, box_loss_weight=1.0,
                 corners_loss_weight=10.0, heading_residual_loss_weight=20.0, size_residual_loss_weight=20.0):
        super().__init__()
        self.box_loss_weight = box_loss_weight
        self.corners_loss_weight = corners_loss_weight
        self.heading_residual_loss_weight = heading_residual_loss_weight
        self.size_residual_loss_weight = size_residual_loss_weight
        self.num_heading_angle_bins = num_heading_angle_bins
        self.num_size_templates = num_size_templates
        self.register_buffer("size_templates", size_templates.view(num_size_templates, 3))
        self.register_buffer("heading_angle_bin_centers",
                             torch.arange(0, 2 * math.pi, 2 * math.pi / num_heading_angle_bins))

    def forward(self, inputs, targets):
        hb, st = targets["heading_bin_id"], targets["size_template_id"]
        center_t = targets["center"]
        rows = torch.arange(inputs["center"].size(0), device=inputs["center"].device)
        bin_width = math.pi / self.num_heading_angle_bins

        seg = TF.cross_entropy(inputs["mask_logits"], targets["mask_logits"])
        head_cls = TF.cross_entropy(inputs["heading_scores"], hb)
        size_cls = TF.cross_entropy(inputs["size_scores"], st)
        center = huber_loss(torch.norm(center_t - inputs["center"], dim=-1), delta=2.0)
        center_reg = huber_loss(torch.norm(center_t - inputs["center_reg"], dim=-1), delta=1.0)

        head_res_n = huber_loss(inputs["heading_residuals_normalized"][rows, hb]
                                - targets["heading_residual"] / bin_width, delta=1.0)
        templates = self.size_templates[st]
        size_res_n = huber_loss(torch.norm(targets["size_residual"] / templates
                                           - inputs["size_residuals_normalized"][rows, st], dim=-1), delta=1.0)

        bin_centers = self.heading_angle_bin_centers[hb]
        pred = get_box_corners_3d(inputs["center"], inputs["heading_residuals"][rows, hb] + bin_centers,
                                  inputs["size_residuals"][rows, st] + templates, with_flip=False)
        gt, gt_flip = get_box_corners_3d(center_t, bin_centers + targets["heading_residual"],
                                         templates + targets["size_residual"], with_flip=True)
        corners = huber_loss(torch.min(torch.norm(pred - gt, dim=1), torch.norm(pred - gt_flip, dim=1)), delta=1.0)

        box = (center + center_reg + head_cls + size_cls + self.heading_residual_loss_weight * head_res_n
               + self.size_residual_loss_weight * size_res_n + self.corners_loss_weight * corners)
        return seg + self.box_loss_weight * box
