"""The four networks BASELINE.json's configs 2-5 name, rebuilt from OUR modules for tests and bench lines
(the reference tree does not travel to the GPU box, and the drop-in `modules` package is what it would import there).

Layer tables follow the reference (cited per class); sub-module names and state_dict keys equal the reference's
(models/s3dis/pvcnn.py, models/s3dis/pvcnnpp.py, models/shapenet/pvcnn.py, models/kitti/frustum/**), which
tests/test_abi_cpu.py checks key by key against the unmodified reference models, so released checkpoints load.
Only the network wiring lives here; every layer is a pvcnn_b200.nn module (PVConv / SharedMLP / PointNet*Module).
"""
import math
import os

import torch
import torch.nn as nn

from . import functional as F
from .nn import PVConv, SharedMLP, PointNetAModule, PointNetSAModule, PointNetFPModule
from .nn.shared_mlp import native_mlp_enabled


def _classify(head, taps, n):
    """cat(taps, dim=1) -> per-point head.  On CUDA the concatenation is written straight into channels-last rows
    (broadcast taps [B,C,1] are expanded on the fly) and the whole head runs on them (pvcnn_b200/mlp.py: cat_cl, head_cl):
    neither the [B, sum C, N] concat nor the repeated cloud feature is materialised."""
    if taps[0].is_cuda and native_mlp_enabled():
        from . import mlp
        if mlp.head_supported(head):
            first = list(head)[0]
            cloud = [t.shape[2] == 1 and n > 1 for t in taps]
            if any(cloud) and not all(cloud) and isinstance(first, SharedMLP) and os.environ.get("PVCNN_B200_HEAD_SPLIT", "1") != "0":
                # Channels that are constant over a cloud (one-hot class vector, max-pooled cloud feature:
                # models/shapenet/pvcnn.py:40-42, models/kitti/frustum/pvcnne.py) never become columns of the point rows: their
                # part of the first layer's weight is applied once per cloud and enters the GEMM epilogue as a per-cloud bias.
                col, pcols, ccols = 0, [], []
                for t, c in zip(taps, cloud):
                    (ccols if c else pcols).append((col, col + t.shape[1]))
                    col += t.shape[1]
                conv0 = list(first.layers)[0]
                gb = mlp.cloud_bias(conv0, [t for t, c in zip(taps, cloud) if c], tuple(ccols))
                rows, lo = mlp.cat_cl([t for t, c in zip(taps, cloud) if not c], n)
                return mlp.head_cl(head, rows, lo, taps[0].shape[0], n, point_cols=tuple(pcols), group_bias=gb)
            rows, lo = mlp.cat_cl(taps, n)
            return mlp.head_cl(head, rows, lo, taps[0].shape[0], n)
    return head(torch.cat([t.expand(-1, -1, n) for t in taps], dim=1))


def _scaled(width, c):
    return int(width * c)


def _point_stack(table, cin, *, with_se=False, normalize=True, eps=0, width=1, vres=1):
    """(channels, repeats, voxel_resolution | None) rows -> PVConv / SharedMLP layers (models/utils.py:48-66)."""
    layers, concat = [], 0
    for channels, repeats, res in table:
        cout = _scaled(width, channels)
        for _ in range(repeats):
            if res is None:
                layers.append(SharedMLP(cin, cout))
            else:
                layers.append(PVConv(cin, cout, 3, int(vres * res), with_se=with_se, normalize=normalize, eps=eps))
            cin = cout
            concat += cout
    return layers, cin, concat


def _head(cin, spec, *, final_linear, per_point, width=1):
    """models/utils.py:15-45: numbers < 1 are dropout rates, the last entry is the (unscaled) output width when
    `final_linear`; per_point -> SharedMLP / Conv1d over [B,C,N], else Linear+BN+ReLU over [B,C]."""
    def block(i, o):
        if per_point:
            return SharedMLP(i, o)
        return nn.Sequential(nn.Linear(i, o), nn.BatchNorm1d(o), nn.ReLU(True))
    layers = []
    for c in spec[:-1]:
        if c < 1:
            layers.append(nn.Dropout(c))
        else:
            layers.append(block(cin, _scaled(width, c)))
            cin = _scaled(width, c)
    if final_linear:
        layers.append(nn.Conv1d(cin, spec[-1], 1) if per_point else nn.Linear(cin, spec[-1]))
        return layers, spec[-1]
    layers.append(block(cin, _scaled(width, spec[-1])))
    return layers, _scaled(width, spec[-1])


class S3DISPVCNN(nn.Module):
    """models/s3dis/pvcnn.py:9-46 (BASELINE config 2: forward, B=16, N=4096)."""
    table = ((64, 1, 32), (64, 2, 16), (128, 1, 16), (1024, 1, None))

    def __init__(self, num_classes=13, extra_feature_channels=6, width_multiplier=1, voxel_resolution_multiplier=1):
        super().__init__()
        self.in_channels = extra_feature_channels + 3
        layers, c_point, c_concat = _point_stack(self.table, self.in_channels, width=width_multiplier,
                                                 vres=voxel_resolution_multiplier)
        self.point_features = nn.ModuleList(layers)
        layers, c_cloud = _head(c_point, [256, 128], final_linear=False, per_point=False, width=width_multiplier)
        self.cloud_features = nn.Sequential(*layers)
        layers, _ = _head(c_concat + c_cloud, [512, 0.3, 256, 0.3, num_classes], final_linear=True, per_point=True,
                          width=width_multiplier)
        self.classifier = nn.Sequential(*layers)

    def forward(self, inputs):
        x = inputs["features"] if isinstance(inputs, dict) else inputs
        coords = x[:, :3, :]
        taps = []
        for layer in self.point_features:
            x, _ = layer((x, coords))
            taps.append(x)
        cloud = self.cloud_features(x.max(dim=-1).values)
        taps.append(cloud.unsqueeze(-1))
        return _classify(self.classifier, taps, coords.size(-1))


class ShapeNetPVCNN(nn.Module):
    """models/shapenet/pvcnn.py:9-42 (BASELINE config 3: train step, width 0.25, B=32, N=2048)."""
    table = ((64, 1, 32), (128, 2, 16), (512, 1, None), (2048, 1, None))

    def __init__(self, num_classes=50, num_shapes=16, extra_feature_channels=3, width_multiplier=1,
                 voxel_resolution_multiplier=1):
        super().__init__()
        self.in_channels = extra_feature_channels + 3
        self.num_shapes = num_shapes
        layers, c_point, c_concat = _point_stack(self.table, self.in_channels, with_se=True, normalize=False,
                                                 width=width_multiplier, vres=voxel_resolution_multiplier)
        self.point_features = nn.ModuleList(layers)
        layers, _ = _head(num_shapes + c_point + c_concat, [256, 0.2, 256, 0.2, 128, num_classes], final_linear=True,
                          per_point=True, width=width_multiplier)
        self.classifier = nn.Sequential(*layers)

    def forward(self, inputs):
        x = inputs[:, :self.in_channels, :]
        n = x.size(-1)
        coords = x[:, :3, :]
        taps = [inputs[:, -self.num_shapes:, :]]
        for layer in self.point_features:
            x, _ = layer((x, coords))
            taps.append(x)
        taps.append(x.max(dim=-1, keepdim=True).values)
        return _classify(self.classifier, taps, n)


class S3DISPVCNN2(nn.Module):
    """models/s3dis/pvcnnpp.py:8-59 (BASELINE config 4: PVCNN++ forward, B=8, N=8192)."""
    sa_table = [((32, 2, 32), (1024, 0.1, 32, (32, 64))), ((64, 3, 16), (256, 0.2, 32, (64, 128))),
                ((128, 3, 8), (64, 0.4, 32, (128, 256))), (None, (16, 0.8, 32, (256, 256, 512)))]
    fp_table = [((256, 256), (256, 1, 8)), ((256, 256), (256, 1, 8)), ((256, 128), (128, 2, 16)),
                ((128, 128, 64), (64, 1, 32))]

    def __init__(self, num_classes=13, extra_feature_channels=6, width_multiplier=1, voxel_resolution_multiplier=1):
        super().__init__()
        w, vr = width_multiplier, voxel_resolution_multiplier
        self.in_channels = extra_feature_channels + 3
        # ---- set abstraction (models/utils.py:69-114)
        cin, extra = self.in_channels, extra_feature_channels
        sa_layers, sa_in = [], []
        for conv_cfg, (centers, radius, neighbors, widths) in self.sa_table:
            sa_in.append(cin)
            stage = []
            if conv_cfg is not None:
                convs, cin, _ = _point_stack([conv_cfg], cin, with_se=True, width=w, vres=vr)
                stage += convs
                extra = cin
            widths = [[_scaled(w, c) for c in x] if isinstance(x, (list, tuple)) else _scaled(w, x) for x in widths]
            if centers is None:
                stage.append(PointNetAModule(in_channels=extra, out_channels=widths, include_coordinates=True))
            else:
                stage.append(PointNetSAModule(num_centers=centers, radius=radius, num_neighbors=neighbors,
                                              in_channels=extra, out_channels=widths, include_coordinates=True))
            cin = extra = stage[-1].out_channels
            sa_layers.append(stage[0] if len(stage) == 1 else nn.Sequential(*stage))
        self.sa_layers = nn.ModuleList(sa_layers)
        # ---- feature propagation (models/utils.py:112-140); only the last FP module sees the raw extra features
        sa_in[0] = extra_feature_channels
        fp_layers = []
        for i, (fp_widths, conv_cfg) in enumerate(self.fp_table):
            widths = tuple(_scaled(w, c) for c in fp_widths)
            stage = [PointNetFPModule(in_channels=cin + sa_in[-1 - i], out_channels=widths)]
            cin = widths[-1]
            if conv_cfg is not None:
                convs, cin, _ = _point_stack([conv_cfg], cin, with_se=True, width=w, vres=vr)
                stage += convs
            fp_layers.append(stage[0] if len(stage) == 1 else nn.Sequential(*stage))
        self.fp_layers = nn.ModuleList(fp_layers)
        layers, _ = _head(cin, [128, 0.5, num_classes], final_linear=True, per_point=True, width=w)
        self.classifier = nn.Sequential(*layers)

    def forward(self, inputs):
        x = inputs["features"] if isinstance(inputs, dict) else inputs
        coords, feats = x[:, :3, :].contiguous(), x
        coord_stack, feat_stack = [], []
        for stage in self.sa_layers:
            feat_stack.append(feats)
            coord_stack.append(coords)
            feats, coords = stage((feats, coords))
        feat_stack[0] = x[:, 3:, :].contiguous()
        for i, stage in enumerate(self.fp_layers):
            feats, coords = stage((coord_stack[-1 - i], coords, feats, feat_stack[-1 - i]))
        return _classify(self.classifier, [feats], feats.size(-1))


class _InstanceSegPVCNN(nn.Module):
    """models/kitti/frustum/segmentation/pointnet.py:9-66 (InstanceSegmentationPVCNN)."""
    point_table = ((64, 2, 16), (64, 1, 12), (128, 1, 12), (1024, 1, None))

    def __init__(self, num_classes=3, extra_feature_channels=1, width_multiplier=1, voxel_resolution_multiplier=1):
        super().__init__()
        self.in_channels = extra_feature_channels + 3
        self.num_classes = num_classes
        layers, c_point, _ = _point_stack(self.point_table, self.in_channels, width=width_multiplier,
                                          vres=voxel_resolution_multiplier)
        self.point_features = nn.Sequential(*layers)
        self.cloud_features = nn.Sequential()
        layers, _ = _head(2 * c_point + num_classes, [512, 256, 128, 128, 0.5, 2], final_linear=True, per_point=True,
                          width=width_multiplier)
        self.classifier = nn.Sequential(*layers)

    def forward(self, inputs):
        x = inputs["features"]
        n = x.size(-1)
        one_hot = inputs["one_hot_vectors"].unsqueeze(-1)
        point, coords = self.point_features((x, x[:, :3, :]))
        cloud, _ = self.cloud_features((point, coords))
        cloud = cloud.max(dim=-1, keepdim=True).values
        return _classify(self.classifier, [one_hot, point, cloud], n)


class _CenterRegression(nn.Module):
    """models/kitti/frustum/center_regression_net.py:9-32."""

    def __init__(self, num_classes=3, width_multiplier=1):
        super().__init__()
        layers, c = _head(3, (128, 128, 256), final_linear=False, per_point=True, width=width_multiplier)
        self.features = nn.Sequential(*layers)
        layers, _ = _head(c + num_classes, [256, 128, 3], final_linear=True, per_point=False, width=width_multiplier)
        self.regression = nn.Sequential(*layers)

    def forward(self, inputs):
        f = self.features(inputs["coords"]).max(dim=-1).values
        return self.regression(torch.cat([f, inputs["one_hot_vectors"]], dim=1))


class _BoxEstimationPointNet(nn.Module):
    """models/kitti/frustum/box_estimation/pointnet.py:9-47."""
    table = ((128, 2, None), (256, 1, None), (512, 1, None))

    def __init__(self, num_classes=3, num_heading_angle_bins=12, num_size_templates=8, width_multiplier=1):
        super().__init__()
        layers, c_point, _ = _point_stack(self.table, 3, normalize=True, eps=1e-15, width=width_multiplier)
        self.features = nn.Sequential(*layers)
        layers, _ = _head(c_point + num_classes, [512, 256, 3 + num_heading_angle_bins * 2 + num_size_templates * 4],
                          final_linear=True, per_point=False, width=width_multiplier)
        self.classifier = nn.Sequential(*layers)

    def forward(self, inputs):
        f, _ = self.features((inputs["coords"], inputs["coords"]))
        return self.classifier(torch.cat([f.max(dim=-1).values, inputs["one_hot_vectors"]], dim=1))


class FrustumPVCNNE(nn.Module):
    """models/kitti/frustum/frustum_net.py:14-100 (BASELINE config 5: end-to-end inference, B=32, N=1024)."""

    def __init__(self, num_classes=3, num_heading_angle_bins=12, num_size_templates=3, num_points_per_object=512,
                 size_templates=None, extra_feature_channels=1, width_multiplier=1, voxel_resolution_multiplier=1):
        super().__init__()
        w = list(width_multiplier) if isinstance(width_multiplier, (list, tuple)) else [width_multiplier] * 3
        self.num_heading_angle_bins = num_heading_angle_bins
        self.num_size_templates = num_size_templates
        self.num_points_per_object = num_points_per_object
        self.inst_seg_net = _InstanceSegPVCNN(num_classes, extra_feature_channels, w[0], voxel_resolution_multiplier)
        self.center_reg_net = _CenterRegression(num_classes, w[1])
        self.box_est_net = _BoxEstimationPointNet(num_classes, num_heading_angle_bins, num_size_templates, w[2])
        if size_templates is None:
            size_templates = torch.ones(num_size_templates, 3)
        self.register_buffer("size_templates", size_templates.view(1, num_size_templates, 3))

    def forward(self, inputs):
        feats, one_hot = inputs["features"], inputs["one_hot_vectors"]
        logits = self.inst_seg_net({"features": feats, "one_hot_vectors": one_hot})
        fg, fg_mean, _ = F.logits_mask(coords=feats[:, :3, :], logits=logits,
                                       num_points_per_object=self.num_points_per_object)
        delta = self.center_reg_net({"coords": fg, "one_hot_vectors": one_hot})
        fg = fg - delta.unsqueeze(-1)
        est = self.box_est_net({"coords": fg, "one_hot_vectors": one_hot})
        nh, ns = self.num_heading_angle_bins, self.num_size_templates
        center, h_score, h_res, s_score, s_res = est.split([3, nh, nh, ns, ns * 3], dim=-1)
        s_res = s_res.view(-1, ns, 3)
        center_reg = fg_mean + delta
        return {"mask_logits": logits, "center_reg": center_reg, "center": center + center_reg,
                "heading_scores": h_score, "heading_residuals_normalized": h_res,
                "heading_residuals": h_res * (math.pi / nh), "size_scores": s_score,
                "size_residuals_normalized": s_res, "size_residuals": s_res * self.size_templates}


def build(config):
    """BASELINE.json config name -> (model, synthetic input factory(batch, generator) -> model input, description)."""
    if config == "s3dis_pvcnn":          # config 2
        return S3DISPVCNN(13, 6), dict(batch=16, points=8 * 512, channels=9, kind="s3dis", mode="eval")
    if config == "shapenet_c0p25_train":  # config 3
        return ShapeNetPVCNN(50, 16, 3, width_multiplier=0.25), dict(batch=32, points=2048, channels=22, kind="shapenet",
                                                                     mode="train")
    if config == "pvcnn2":               # config 4
        return S3DISPVCNN2(13, 6), dict(batch=8, points=8192, channels=9, kind="s3dis", mode="eval")
    if config == "frustum_pvcnne":       # config 5
        return FrustumPVCNNE(), dict(batch=32, points=1024, channels=4, kind="frustum", mode="eval")
    raise ValueError(config)


def synthetic_input(spec, generator, device="cpu", batch=None):
    """SURVEY.md 8d synthetic inputs: S3DIS [xyz block, rgb, xyz room], ShapeNet [xyz, normals, one-hot(16)],
    frustum {'features': [xyz, intensity], 'one_hot_vectors'}."""
    b = batch or spec["batch"]
    n = spec["points"]
    g = generator
    if spec["kind"] == "s3dis":
        xyz = torch.rand(b, 3, n, generator=g) * torch.tensor([1.5, 1.5, 3.0]).view(1, 3, 1)
        x = torch.cat([xyz, torch.rand(b, 6, n, generator=g)], dim=1)
        return x.to(device)
    if spec["kind"] == "shapenet":
        xyz = torch.randn(b, 3, n, generator=g)
        xyz = xyz / xyz.norm(dim=1, keepdim=True).clamp_min(1e-6) * torch.rand(b, 1, n, generator=g).pow(1 / 3)
        normals = torch.randn(b, 3, n, generator=g)
        normals = normals / normals.norm(dim=1, keepdim=True).clamp_min(1e-6)
        shape = torch.randint(0, 16, (b,), generator=g)
        one_hot = torch.zeros(b, 16, n)
        one_hot[torch.arange(b), shape] = 1.0
        return torch.cat([xyz, normals, one_hot], dim=1).to(device)
    if spec["kind"] == "frustum":
        xyz = torch.randn(b, 3, n, generator=g) * torch.tensor([1.0, 0.6, 2.5]).view(1, 3, 1)
        feats = torch.cat([xyz, torch.rand(b, 1, n, generator=g)], dim=1)
        one_hot = torch.zeros(b, 3)
        one_hot[torch.arange(b), torch.randint(0, 3, (b,), generator=g)] = 1.0
        return {"features": feats.to(device), "one_hot_vectors": one_hot.to(device)}
    raise ValueError(spec["kind"])
