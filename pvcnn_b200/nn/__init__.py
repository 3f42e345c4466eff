"""`modules` surface of the reference (modules/__init__.py:1-8)."""
from .ball_query import BallQuery
from .frustum import FrustumPointNetLoss, get_box_corners_3d
from .loss import KLLoss
from .pointnet import PointNetAModule, PointNetSAModule, PointNetFPModule
from .pvconv import PVConv
from .se import SE3d
from .shared_mlp import SharedMLP
from .voxelization import Voxelization

__all__ = ["BallQuery", "FrustumPointNetLoss", "KLLoss", "PointNetAModule", "PointNetSAModule",
           "PointNetFPModule", "PVConv", "SE3d", "SharedMLP", "Voxelization", "get_box_corners_3d"]
