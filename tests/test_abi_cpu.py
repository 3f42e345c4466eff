"""CPU tests of the boundary: the C-ABI library loads and exports every symbol the header
declares; the drop-in `modules` package exposes the reference's names; the reference's own
models build on top of it (when /root/reference is present)."""
import ctypes
import os
import sys

import pytest

from pvcnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    syms = _lib.header_symbols()
    assert len(syms) >= 16
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert lib.pvcnn_abi_version() == 1


def test_modules_surface():
    import modules
    import modules.functional as F
    from modules.frustum import get_box_corners_3d, FrustumPointNetLoss  # noqa: F401
    for name in ["BallQuery", "FrustumPointNetLoss", "KLLoss", "PointNetAModule", "PointNetSAModule",
                 "PointNetFPModule", "PVConv", "SE3d", "SharedMLP", "Voxelization"]:
        assert hasattr(modules, name)
    for name in ["ball_query", "trilinear_devoxelize", "grouping", "nearest_neighbor_interpolate", "kl_loss",
                 "huber_loss", "gather", "furthest_point_sample", "logits_mask", "avg_voxelize"]:
        assert hasattr(F, name)


def test_pvconv_state_dict_matches_reference_layout():
    import modules
    m = modules.PVConv(64, 64, 3, 32, with_se=True)
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert sd["voxel_layers.0.weight"] == (64, 64, 3, 3, 3)
    assert sd["voxel_layers.3.weight"] == (64, 64, 3, 3, 3)
    assert sd["voxel_layers.1.running_var"] == (64,)
    assert sd["voxel_layers.6.fc.0.weight"] == (8, 64) and sd["voxel_layers.6.fc.2.weight"] == (64, 8)
    assert sd["point_features.layers.0.weight"] == (64, 64, 1)
    assert "point_features.layers.1.num_batches_tracked" in sd


def test_cpu_tensors_fail_loudly():
    import torch
    import modules.functional as F
    with pytest.raises(RuntimeError):
        F.avg_voxelize(torch.zeros(1, 2, 8), torch.zeros(1, 3, 8, dtype=torch.int32), 2)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present")
def test_reference_models_build_on_our_modules():
    """The reference's models/ (unmodified, imported from /root/reference) must compose our
    `modules` package: same constructor signatures and state_dict layout."""
    import importlib
    import torch
    added = "/root/reference"
    # our repo root must win for `modules`; the reference only supplies `models`
    sys.path.append(added)
    try:
        for mod in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[mod]
        import modules
        assert os.path.dirname(modules.__file__).startswith(ROOT)
        s3dis = importlib.import_module("models.s3dis")
        shapenet = importlib.import_module("models.shapenet")
        kitti = importlib.import_module("models.kitti.frustum")
        net = s3dis.PVCNN(num_classes=13, extra_feature_channels=6)
        assert sum(p.numel() for p in net.parameters()) == 2572493
        net2 = s3dis.PVCNN2(num_classes=13, extra_feature_channels=6)
        assert sum(p.numel() for p in net2.parameters()) == 13709837
        net3 = shapenet.PVCNN(num_classes=50, num_shapes=16, extra_feature_channels=3, width_multiplier=0.25)
        assert sum(p.numel() for p in net3.parameters()) == 268898
        assert kitti.FrustumPVCNNE is not None
        assert isinstance(net.point_features[0], modules.PVConv)
        # pvcnn_b200.zoo rebuilds the same four networks from our modules (what tests / bench use on the GPU box, where
        # the reference tree is absent): state_dict keys AND shapes must equal the unmodified reference models'
        from pvcnn_b200 import zoo
        ones = torch.ones(3, 3)
        ref_frustum = kitti.FrustumPVCNNE(num_classes=3, num_heading_angle_bins=12, num_size_templates=3,
                                          num_points_per_object=512, size_templates=ones)
        for ours, ref in ((zoo.S3DISPVCNN(13, 6), net), (zoo.S3DISPVCNN2(13, 6), net2),
                          (zoo.ShapeNetPVCNN(50, 16, 3, width_multiplier=0.25), net3),
                          (zoo.FrustumPVCNNE(size_templates=ones), ref_frustum)):
            so, sr = ours.state_dict(), ref.state_dict()
            assert list(so.keys()) == list(sr.keys()), type(ours).__name__
            assert all(so[k].shape == sr[k].shape for k in so), type(ours).__name__
            ours.load_state_dict(sr)   # a reference checkpoint loads unchanged
    finally:
        sys.path.remove(added)


def test_invalid_arguments_are_rejected_before_any_gpu_work():
    """Every launcher validates sizes / pointers first and returns a non-zero status (the reference would launch and
    `exit(-1)`, cuda_utils.cuh:28-37).  No CUDA call is reached, so this runs on the GPU-less box."""
    import ctypes
    from pvcnn_b200 import _lib
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    i = ctypes.c_int
    f = ctypes.c_float
    ll, ull = ctypes.c_longlong, ctypes.c_ulonglong
    cases = {
        "pvcnn_avg_voxelize": [i(0), i(4), i(8), i(2), i(4), i(8), null, null, null, null, null, null],
        "pvcnn_avg_voxelize_grad": [i(1), i(4), i(8), i(8), null, null, null, null, null],
        "pvcnn_trilinear_devoxelize": [i(1), i(4), i(8), i(2), i(4), i(8), i(1), null, null, null, null, null, null],
        "pvcnn_trilinear_devoxelize_grad": [i(1), i(4), i(8), i(8), null, null, null, null, null],
        "pvcnn_ball_query": [i(1), i(8), i(0), f(0.1), i(4), null, null, null, null],
        "pvcnn_grouping": [i(1), i(4), i(8), i(2), i(4), null, null, null, null],
        "pvcnn_grouping_grad": [i(1), i(4), i(8), i(2), i(4), null, null, null, null],
        "pvcnn_group_concat": [i(1), i(4), i(8), i(2), i(4), null, null, null, null, null, null],
        "pvcnn_group_concat_grad": [i(1), i(4), i(8), i(2), i(4), null, null, null, null, null, null],
        "pvcnn_gather_features": [i(1), i(4), i(8), i(2), null, null, null, null],
        "pvcnn_gather_features_grad": [i(1), i(4), i(8), i(2), null, null, null, null],
        "pvcnn_furthest_point_sampling": [i(1), i(8), i(2), null, null, null, null],
        "pvcnn_pvconv_forward": [null, null, null, null, null, null, null],
        "pvcnn_pvconv_backward": [null, null, null, null, null, null, null],
        # test-time voting (csrc/eval_voting.cu): (b, nv, seed, first_window, num_points, indices, stream) ...
        "pvcnn_vote_indices": [i(0), i(8), ull(1), i(0), null, null, null],
        "pvcnn_window_indices": [i(1), i(0), ull(1), i(0), null, null, null],
        "pvcnn_vote_gather": [i(1), i(3), i(8), i(1), i(8), i(1), null, null, null, null, null, null],
        "pvcnn_softmax_max": [i(1), i(4), i(8), i(2), i(2), null, null, null, null],
        "pvcnn_vote_reset": [ll(0), null, null, null],
        "pvcnn_vote_merge": [i(1), i(8), i(1), ll(10), ctypes.c_uint(0), null, null, null, null, null, null, null],
        "pvcnn_vote_confidences": [ll(5), null, null, null],
        "pvcnn_vote_stats": [ll(5), i(5000), i(1), null, null, null, null],
    }
    bad_arg = 100001   # PVCNN_E_BADARG (include/pvcnn_b200.h:38); a CUDA failure would surface as a cudaError_t (< 1000)
    for name, args in cases.items():
        assert getattr(lib, name)(*args) == bad_arg, name


def test_voting_host_api_has_no_cpu_path():
    """pvcnn_b200.evaluate mirrors the reference's evaluation loops on the device only: host tensors raise"""
    import torch
    from pvcnn_b200 import _lib, evaluate
    with pytest.raises(_lib.PvcnnError):
        evaluate.softmax_max(torch.zeros(1, 3, 8))
    with pytest.raises(RuntimeError):
        evaluate.softmax_max(torch.zeros(3, 8))
    with pytest.raises(RuntimeError):
        evaluate.vote_inputs(torch.zeros(1, 8, 3), torch.zeros(1, 6, dtype=torch.int32), 4)   # nv % num_points != 0


def test_workspace_size_queries_are_host_only_and_consistent():
    """pvcnn_pvconv_*_floats / _ints size the caller-owned workspace; pure host arithmetic."""
    import ctypes
    from pvcnn_b200 import _lib, fused
    lib = _lib.load()
    for fn in ("pvcnn_pvconv_wprep_floats", "pvcnn_pvconv_partials_floats", "pvcnn_pvconv_sparse_ints"):
        getattr(lib, fn).restype = ctypes.c_longlong

    def desc(b, n, cin, cout, r, npass=3):
        return fused.Desc(b, n, cin, cout, r, 1, 0.0, 1, npass, 1e-4, 1e-5, 0.1, 0.1, 0)

    d = desc(16, 4096, 64, 64, 32)
    w = lib.pvcnn_pvconv_wprep_floats(ctypes.byref(d))
    # 2 x (hi, lo) x [27 (w1) + 27 (w2) + 1 (wp)] x 64 x 64, forward and data-gradient operand copies
    assert w == 2 * 2 * (27 + 27 + 1) * 64 * 64
    assert lib.pvcnn_pvconv_needs_grid_lo(ctypes.byref(d)) == 1
    assert lib.pvcnn_pvconv_needs_grid_lo(ctypes.byref(desc(16, 4096, 64, 64, 32, npass=1))) == 0
    small, big = desc(2, 1024, 16, 16, 8), desc(4, 2048, 9, 64, 16)
    for fn in ("pvcnn_pvconv_wprep_floats", "pvcnn_pvconv_partials_floats", "pvcnn_pvconv_sparse_ints"):
        a, b_ = getattr(lib, fn)(ctypes.byref(small)), getattr(lib, fn)(ctypes.byref(big))
        assert 0 < a <= b_, fn
    # activity lists: 8 counters + one int4 per unit / k-tile + flags + 4 x 27 x C class sums
    units, kt = 16 * 16 * 8, 16 * 32 * 32 * 1
    assert lib.pvcnn_pvconv_sparse_ints(ctypes.byref(d)) >= 8 + 4 * 4 * units + 2 * 4 * kt + 4 * 27 * 64
