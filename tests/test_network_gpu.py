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


def _l2_rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _quantile_err(a, b, q):
    """q-quantile of |a - b| relative to the rms of b."""
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.quantile(np.abs(a - b), q) / max(np.sqrt(np.mean(b * b)), 1e-30))


def _make_case(seed=50, b=4, n=2048):
    g = rng(seed)
    x = np.concatenate([s3dis_like_coords(g, b, n), g.random((b, 6, n), dtype=np.float32)], axis=1)
    labels = torch.from_numpy(g.integers(0, 13, size=(b, n))).cuda()
    torch.manual_seed(3)
    net = MiniPVCNN().cuda().train()
    return x, labels, net, {k: v.clone() for k, v in net.state_dict().items()}


def test_mini_network_fused_vs_composed(monkeypatch):
    """Whole-network comparison.  Logits / loss / running statistics are compared tightly.  Gradients are not well
    conditioned through three train-mode BatchNorm + ReLU stages (a 1e-6 difference in a block's input flips ReLU masks
    and is amplified ~100x by the BN-backward cancellation -- measured: each mode matches an fp64 evaluation on ITS OWN
    inputs to 1e-6, tools/net_block2b.py), so they are compared in the L2 norm here and block by block on identical
    inputs, tightly, in test_mini_network_blocks_in_context."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x, labels, net, state = _make_case()
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
    assert _l2_rel(res["fused"][2], res["composed"][2]) < 2e-2
    for k, gref in res["composed"][3].items():
        if np.abs(gref).max() < 1e-7:   # biases in front of a train-mode BatchNorm: zero gradient, noise only
            continue
        assert np.isfinite(res["fused"][3][k]).all(), k
        assert _l2_rel(res["fused"][3][k], gref) < 2e-2, k
    for k, bref in res["composed"][4].items():   # BatchNorm running statistics follow torch's update rule
        assert np.abs(res["fused"][4][k] - bref).max() < 1e-4 * max(1.0, np.abs(bref).max()), k


def test_mini_network_blocks_in_context(monkeypatch):
    """Every fused block, run INSIDE the network (its saved activations must survive the other blocks' forward and
    backward passes; scratch buffers are shared), against the same block run stand-alone in composed mode on the very
    same input / coordinates / output gradient."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x, labels, net, state = _make_case(seed=52)
    monkeypatch.setenv("PVCNN_B200_PVCONV", "fused")
    cap = [dict() for _ in range(3)]
    hooks = []
    for k in range(3):
        def _hook(m, i, o, k=k):
            o[0].retain_grad()
            cap[k].update(fin=i[0][0], coords=i[0][1].detach().clone(), out=o[0])
        hooks.append(net.blocks[k].register_forward_hook(_hook))
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    loss = nn.functional.cross_entropy(net(xt), labels)
    loss.backward(retain_graph=True)
    for h in hooks:
        h.remove()
    fused = []
    for k in range(3):   # first block first: its buffers are the oldest
        blk, c = net.blocks[k], cap[k]
        gout = c["out"].grad.detach().clone()
        pgrads = {n_: p.grad.detach().clone() for n_, p in blk.named_parameters()}
        gin, = torch.autograd.grad(c["out"], c["fin"], gout, retain_graph=True)
        fused.append(dict(out=c["out"].detach().clone(), gout=gout, gin=gin.clone(), pgrads=pgrads,
                          fin=c["fin"].detach().clone(), coords=c["coords"]))
    monkeypatch.setenv("PVCNN_B200_PVCONV", "composed")
    net.load_state_dict(state)
    for k in range(3):
        blk, f = net.blocks[k], fused[k]
        blk.zero_grad(set_to_none=True)
        fin = f["fin"].clone().requires_grad_(True)
        out, _ = blk((fin, f["coords"]))
        out.backward(f["gout"])
        assert rel_err(f["out"].cpu().numpy(), out.detach().cpu().numpy()) < 2e-5, k
        # Gradients: an activation within fp32 rounding of zero can land on different sides of a (Leaky)ReLU in the two
        # implementations (expected a few times per 10 runs at this size); that changes one voxel-channel's gradient 10x
        # and everything downstream of it.  Element-wise agreement is therefore asserted for 99 % of the input gradient,
        # and the L2 error of every gradient must stay at the single-flip level; a lost / overwritten buffer gives O(1).
        assert _quantile_err(f["gin"].cpu().numpy(), fin.grad.cpu().numpy(), 0.99) < 5e-4, k
        assert _l2_rel(f["gin"].cpu().numpy(), fin.grad.cpu().numpy()) < 2e-2, k
        for n_, p in blk.named_parameters():
            gref = p.grad.cpu().numpy()
            if n_ in ("voxel_layers.0.bias", "voxel_layers.3.bias", "point_features.layers.0.bias"):
                continue   # conv biases in front of a train-mode BatchNorm: exactly-zero gradient, rounding noise only
            assert _l2_rel(f["pgrads"][n_].cpu().numpy(), gref) < 2e-2, (k, n_)


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
    assert torch.allclose(y1, y2, rtol=1e-4, atol=1e-5)  # scatter atomics are order-nondeterministic: not bit-equal
    assert torch.isfinite(y1).all()
