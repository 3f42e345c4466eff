"""A small PVCNN-style network assembled from the drop-in modules (the reference's models/ cannot travel to the GPU
box): several PVConv blocks chained with different widths / resolutions / SE, a global max-pool branch and a classifier.
Checks that the fused blocks compose (shared scratch buffers, saved activations per block) by comparing the whole
network, forward and backward, against the same network run block-by-block in 'composed' mode (stand-alone sm_100a ops
around torch's dense layers with TF32 disabled)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import modules
from util import rng, s3dis_like_coords, rel_err

pytestmark = pytest.mark.gpu


class MiniPVCNN(nn.Module):
    def __init__(self, num_classes=13, in_channels=9):
        super().__init__()
        self.blocks = nn.ModuleList([
            modules.PVConv(in_channels, 32, 3, 16),
            modules.PVConv(32, 64, 3, 16, with_se=True),
            modules.PVConv(64, 64, 3, 8),
            modules.SharedMLP(64, 128),
        ])
        self.cloud = nn.Sequential(nn.Linear(128, 64), nn.BatchNorm1d(64), nn.ReLU(True))
        self.classifier = nn.Sequential(modules.SharedMLP(32 + 64 + 64 + 128 + 64, 64), nn.Conv1d(64, num_classes, 1))

    def forward(self, x):
        coords = x[:, :3, :]
        feats, outs = x, []
        for blk in self.blocks:
            feats, _ = blk((feats, coords))
            outs.append(feats)
        pooled = self.cloud(feats.max(dim=-1).values)
        outs.append(pooled.unsqueeze(-1).repeat(1, 1, x.size(-1)))
        return self.classifier(torch.cat(outs, dim=1))


def test_mini_network_fused_vs_composed(monkeypatch):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = rng(50)
    b, n = 4, 2048
    x = np.concatenate([s3dis_like_coords(g, b, n), g.random((b, 6, n), dtype=np.float32)], axis=1)
    labels = torch.from_numpy(g.integers(0, 13, size=(b, n))).cuda()
    torch.manual_seed(3)
    net = MiniPVCNN().cuda().train()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    res = {}
    for mode in ("composed", "fused"):
        monkeypatch.setenv("PVCNN_B200_PVCONV", mode)
        net.load_state_dict(state)
        net.zero_grad(set_to_none=True)
        xt = torch.from_numpy(x).cuda().requires_grad_(True)
        logits = net(xt)
        loss = nn.functional.cross_entropy(logits, labels)
        loss.backward()
        res[mode] = (logits.detach().cpu().numpy(), float(loss), xt.grad.cpu().numpy(),
                     {k: p.grad.cpu().numpy() for k, p in net.named_parameters()},
                     {k: v.detach().cpu().numpy() for k, v in net.state_dict().items() if "running" in k})
    assert rel_err(res["fused"][0], res["composed"][0]) < 5e-5
    assert abs(res["fused"][1] - res["composed"][1]) < 1e-5
    assert rel_err(res["fused"][2], res["composed"][2]) < 2e-4
    for k, gref in res["composed"][3].items():
        if np.abs(gref).max() < 1e-7:   # biases in front of a train-mode BatchNorm: zero gradient, noise only
            continue
        assert rel_err(res["fused"][3][k], gref) < 5e-4, k
    for k, bref in res["composed"][4].items():   # BatchNorm running statistics follow torch's update rule
        assert np.abs(res["fused"][4][k] - bref).max() < 1e-4 * max(1.0, np.abs(bref).max()), k


def test_mini_network_eval_and_state_dict_roundtrip(monkeypatch):
    monkeypatch.setenv("PVCNN_B200_PVCONV", "fused")
    g = rng(51)
    b, n = 2, 1024
    x = torch.from_numpy(np.concatenate([s3dis_like_coords(g, b, n), g.random((b, 6, n), dtype=np.float32)], axis=1)).cuda()
    torch.manual_seed(4)
    net = MiniPVCNN().cuda()
    with torch.no_grad():
        net.train(); net(x); net.eval()
        y1 = net(x)
        net2 = MiniPVCNN().cuda().eval()
        net2.load_state_dict(net.state_dict())
        y2 = net2(x)
    assert torch.equal(y1, y2)
    assert torch.isfinite(y1).all()
