"""Drop-in `modules` package: the import name the reference's models/, configs/ and meters/
use (models/utils.py:5, models/kitti/frustum/frustum_net.py:6, meters/kitti/frustum.py:4).
Everything is served by pvcnn_b200 (sm_100a library behind include/pvcnn_b200.h)."""
from pvcnn_b200.nn import (BallQuery, FrustumPointNetLoss, KLLoss, PointNetAModule, PointNetSAModule,
                           PointNetFPModule, PVConv, SE3d, SharedMLP, Voxelization)

__all__ = ["BallQuery", "FrustumPointNetLoss", "KLLoss", "PointNetAModule", "PointNetSAModule",
           "PointNetFPModule", "PVConv", "SE3d", "SharedMLP", "Voxelization"]
