"""PointNet++ style modules (BallQuery / SA / FP) on the GPU: composition of the stand-alone sm_100a ops."""
import numpy as np
import pytest
import torch

import oracle
import modules
from util import rng, rel_err

pytestmark = pytest.mark.gpu


def test_ball_query_module_matches_oracle():
    g = rng(40)
    b, n, m, c, u = 2, 1024, 128, 6, 16
    pts = g.random((b, 3, n), dtype=np.float32)
    feats = g.standard_normal((b, c, n), dtype=np.float32)
    fidx = oracle.furthest_point_sampling(pts, m)
    centers = oracle.gather(pts, fidx)
    bq = modules.BallQuery(0.2, u, include_coordinates=True)
    out = bq(torch.from_numpy(pts).cuda(), torch.from_numpy(centers).cuda(), torch.from_numpy(feats).cuda())
    idx = oracle.ball_query(centers, pts, 0.2, u)
    rel = oracle.grouping(pts, idx) - centers[:, :, :, None]
    want = np.concatenate([rel, oracle.grouping(feats, idx)], axis=1)
    assert out.shape == (b, 3 + c, m, u)
    assert np.array_equal(out.cpu().numpy(), want)


def test_sa_fp_modules_forward_backward():
    torch.manual_seed(0)
    g = rng(41)
    b, n, c = 2, 2048, 16
    feats = torch.from_numpy(g.standard_normal((b, c, n), dtype=np.float32)).cuda().requires_grad_(True)
    coords = torch.from_numpy(g.random((b, 3, n), dtype=np.float32)).cuda()
    sa = modules.PointNetSAModule(num_centers=256, radius=[0.1, 0.2], num_neighbors=[16, 32], in_channels=c,
                                  out_channels=[[16, 32], [16, 32]]).cuda()
    fp = modules.PointNetFPModule(in_channels=64 + c, out_channels=[32, 24]).cuda()
    sa_feats, centers = sa((feats, coords))
    assert sa_feats.shape == (b, 64, 256) and centers.shape == (b, 3, 256)
    # centres are the FPS picks of the oracle (index-exact)
    want_c = oracle.gather(coords.cpu().numpy(), oracle.furthest_point_sampling(coords.cpu().numpy(), 256))
    assert np.array_equal(centers.detach().cpu().numpy(), want_c)
    out, _ = fp((coords, centers, sa_feats, feats))
    assert out.shape == (b, 24, n)
    out.square().mean().backward()
    assert feats.grad is not None and torch.isfinite(feats.grad).all() and feats.grad.abs().sum() > 0
    for p in list(sa.parameters()) + list(fp.parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all()


def test_pointnet_a_module_and_logits_mask():
    torch.manual_seed(1)
    g = rng(42)
    b, n, c = 3, 512, 8
    feats = torch.from_numpy(g.standard_normal((b, c, n), dtype=np.float32)).cuda()
    coords = torch.from_numpy(g.random((b, 3, n), dtype=np.float32)).cuda()
    a = modules.PointNetAModule(c, [16, 32]).cuda()
    out, origin = a((feats, coords))
    assert out.shape == (b, 32, 1) and origin.shape == (b, 3, 1)
    import modules.functional as F
    logits = torch.from_numpy(g.standard_normal((b, 2, n), dtype=np.float32)).cuda()
    np.random.seed(0)
    sel, mean, mask = F.logits_mask(coords, logits, 128)
    assert sel.shape == (b, 3, 128) and mean.shape == (b, 3) and mask.shape == (b, n)
    m0 = (logits[:, 0] < logits[:, 1])
    assert torch.equal(mask, m0)


@pytest.mark.parametrize("b,n,k", [(32, 1024, 512), (5, 777, 128), (3, 300, 700), (2, 4096, 2048)])
def test_logits_mask_device_sampling_matches_oracle(b, n, k):
    """the per-sample host loop of modules/functional/sampling.py:66-82 runs on the device.  Picks are
    bit-identical to the numpy restatement with the same counter-based generator (oracle.logits_mask_sample); mask and
    masked mean follow the reference's tensor program; foreground fractions cover nc = 0, nc < k and nc >= k."""
    import modules.functional as F
    g = rng(43)
    coords = g.random((b, 3, n), dtype=np.float32)
    logits = g.standard_normal((b, 2, n), dtype=np.float32)
    frac = np.linspace(0.0, 1.0, b)          # sample 0 has no foreground point at all
    logits[:, 1] = np.where(g.random((b, n)) < frac[:, None], logits[:, 0] + 1.0, logits[:, 0] - 1.0)
    np.random.seed(7)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    np.random.seed(7)
    ct, lt = torch.from_numpy(coords).cuda(), torch.from_numpy(logits).cuda()
    sel, mean, mask = F.logits_mask(ct, lt, k)
    m0 = logits[:, 0] < logits[:, 1]
    assert np.array_equal(mask.cpu().numpy(), m0)
    picks = oracle.logits_mask_sample(m0, k, seed)
    masked = coords * m0[:, None, :]
    mean0 = masked.sum(-1, dtype=np.float64) / np.maximum(m0.sum(-1, keepdims=True), 1)
    assert rel_err(mean.cpu().numpy(), mean0) < 1e-5
    want = np.take_along_axis(masked - mean.cpu().numpy()[:, :, None], picks[:, None, :].astype(np.int64).repeat(3, 1), axis=2)
    assert np.array_equal(sel.cpu().numpy(), want.astype(np.float32))
    # distribution-free properties of the reference's scheme
    for i in range(b):
        cand = set(np.nonzero(m0[i])[0].tolist())
        if not cand:
            assert (picks[i] == 0).all()
            continue
        assert set(picks[i].tolist()) <= cand
        counts = np.bincount(picks[i], minlength=n)[sorted(cand)]
        if len(cand) >= k:
            assert counts.max() == 1
        else:
            assert counts.min() >= k // len(cand) and counts.max() <= k // len(cand) + 1


def test_logits_mask_numpy_mode_keeps_reference_rng_sequence(monkeypatch):
    import modules.functional as F
    monkeypatch.setenv("PVCNN_B200_LOGITS_MASK", "numpy")
    g = rng(44)
    b, n, k = 4, 512, 128
    coords = torch.from_numpy(g.random((b, 3, n), dtype=np.float32)).cuda()
    logits = torch.from_numpy(g.standard_normal((b, 2, n), dtype=np.float32)).cuda()
    np.random.seed(3)
    sel, mean, mask = F.logits_mask(coords, logits, k)
    np.random.seed(3)
    m0 = mask.cpu().numpy()
    masked = coords.cpu().numpy() * m0[:, None, :]
    for i in range(b):   # the reference's call sequence (sampling.py:69-82)
        cand = np.nonzero(m0[i])[0]
        ch = np.random.choice(cand.size, k, replace=False) if cand.size >= k else None
        assert ch is not None
        want = (masked[i] - mean[i].cpu().numpy()[:, None])[:, cand[ch]]
        assert np.array_equal(sel[i].cpu().numpy(), want.astype(np.float32))
