"""`modules.frustum` alias (reference: modules/frustum.py)."""
from pvcnn_b200.nn.frustum import FrustumPointNetLoss, get_box_corners_3d

__all__ = ["FrustumPointNetLoss", "get_box_corners_3d"]
