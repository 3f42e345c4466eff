"""GPU parity tests of the test-time voting kernels (csrc/eval_voting.cu through the C ABI, pvcnn_b200/evaluate.py).

The merge / statistics kernels are compared with the REFERENCE's own outputs: tests/golden/ref_voting_golden.npz holds
what the unmodified numba functions of evaluate/s3dis/eval.py and evaluate/shapenet/eval.py produced (tests/golden/
make_voting_golden.py); integer results must be bit-exact.  Index generation is compared bit for bit with the oracle's
restatement of the counter-based generator, the gather with the literal numpy tiling of eval.py:157-171, softmax-max
with the reference's torch calls (1e-5 relative, the north-star tolerance for fp32)."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_voting as ev

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_voting_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(G)


def _case(gold, name):
    return {k.split(".", 1)[1]: gold[k] for k in gold.files if k.startswith(name + ".")}


def _cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("nv,first", [(4096, 0), (3 * 4096, 10), (1000, 3), (20 * 4096, 1), (2, 0), (1, 5)])
def test_vote_indices_bit_exact_vs_oracle(nv, first):
    from pvcnn_b200 import evaluate as E
    num_points = np.array([700, 4096, 1, 5000, 0, 4097, 33], np.int32)
    got = E.vote_indices(num_points, nv, 0x1234567890ABCDEF, first).cpu().numpy()
    assert np.array_equal(got, ev.vote_indices(num_points, nv, 0x1234567890ABCDEF, first))
    for w, n in enumerate(num_points):   # the multiset of eval.py:161-163, whatever the generator
        if n > 0:
            expect = np.bincount(np.tile(np.arange(n), -(-nv // n))[:nv], minlength=n)
            assert np.array_equal(np.sort(np.bincount(got[w], minlength=n)), np.sort(expect))


@pytest.mark.parametrize("k", [4096, 1024, 7])
def test_window_indices_bit_exact_vs_oracle(k):
    from pvcnn_b200 import evaluate as E
    num_points = np.array([5000, 4096, 100, 1, 0, 8191, 65537], np.int32)
    got = E.window_indices(num_points, k, 77, 2).cpu().numpy()
    assert np.array_equal(got, ev.window_indices(num_points, k, 77, 2))
    for w, n in enumerate(num_points):
        if n >= k:
            assert np.unique(got[w]).size == k and got[w].max() < n      # np.random.choice(replace=False)
        elif n > 0:
            assert got[w].max() < n and got[w].min() >= 0


def test_vote_gather_equals_literal_numpy_tiling():
    from pvcnn_b200 import evaluate as E
    g = np.random.default_rng(3)
    b, p, ch, npo, extra = 5, 777, 9, 256, 6
    data = g.standard_normal((b, p, ch)).astype(np.float32)
    labels = g.integers(0, 13, size=(b, p)).astype(np.int64)
    npts = np.array([777, 300, 1, 512, 700], np.int32)
    idx = ev.vote_indices(npts, extra * npo, 5)
    out, lab = E.vote_inputs(_cuda(data), _cuda(idx), npo, labels=_cuda(labels))
    assert np.array_equal(out.cpu().numpy(), ev.vote_inputs(data, idx, npo))                  # eval.py:157-171
    assert np.array_equal(lab.cpu().numpy(), np.take_along_axis(labels, idx.astype(np.int64), 1))
    # shapenet layout: point_set [ch, n] (shapenet eval.py:158-160)
    ps = g.standard_normal((ch, p)).astype(np.float32)
    so = E.vote_inputs(_cuda(ps)[None], _cuda(idx[:1]), npo, channels_last=False)
    assert np.array_equal(so.cpu().numpy(), ev.shape_inputs(ps, idx[0], npo))


@pytest.mark.parametrize("b,c,n,c0,c1", [(4, 13, 4096, 0, 13), (3, 50, 2048, 12, 16), (2, 2, 1000, 0, 2), (1, 1, 5, 0, 1),
                                         (2, 50, 333, 47, 50)])
def test_softmax_max_vs_reference_torch_calls(b, c, n, c0, c1):
    from pvcnn_b200 import evaluate as E
    g = np.random.default_rng(b * 100 + c)
    logits = (g.standard_normal((b, c, n)) * 4).astype(np.float32)
    conf, pred = E.softmax_max(_cuda(logits), c0, c1)
    conf, pred = conf.cpu().numpy(), pred.cpu().numpy()
    rconf, rpred = ev.softmax_max(logits, c0, c1)
    assert np.abs(conf - rconf).max() <= 1e-5 * np.abs(rconf).max()
    sm = torch.softmax(torch.from_numpy(logits).double(), 1).numpy()[:, c0:c1]
    top = np.sort(sm, axis=1)
    gap = top[:, -1] - top[:, -2] if c1 - c0 > 1 else np.ones((b, n))
    clear = gap > 1e-5 * top[:, -1]      # the two best classes are not within rounding of each other
    assert np.array_equal(pred[clear], rpred[clear])
    assert pred.min() >= c0 and pred.max() < c1
    assert clear.mean() > 0.98


def test_softmax_max_ties_keep_first_class():
    from pvcnn_b200 import evaluate as E
    logits = torch.zeros(1, 6, 64, device="cuda")
    logits[0, 2] = 1.0
    logits[0, 4] = 1.0
    conf, pred = E.softmax_max(logits)
    assert (pred == 2).all()                              # torch.max: first maximal value
    conf2, pred2 = E.softmax_max(logits, 3, 6)
    assert (pred2 == 4).all() and torch.equal(conf, conf2)


@pytest.mark.parametrize("name", ["s3dis_ties", "s3dis_dense", "s3dis_sparse"])
def test_scene_merge_and_stats_equal_reference_numba_outputs(gold, name):
    """update_scene_predictions + update_stats of the unmodified reference (golden) vs the device merge, batch by batch"""
    from pvcnn_b200 import evaluate as E
    c = _case(gold, name)
    num_windows, nv = c["conf"].shape
    bs = int(c["batch_size"])
    votes = E.SceneVotes(c["out_conf"].size)
    mapping = _cuda(c["mapping"], torch.int32)
    for lo in range(0, num_windows, bs):
        hi = min(lo + bs, num_windows)
        votes.update(_cuda(c["conf"][lo:hi]), _cuda(c["pred"][lo:hi], torch.int32), _cuda(c["idx"][lo:hi], torch.int32),
                     mapping[lo:hi])
    assert np.array_equal(votes.predictions.cpu().numpy().astype(np.int64), c["out_pred"])
    assert np.array_equal(votes.confidences.cpu().numpy(), c["out_conf"])
    stats = votes.stats(c["gt"], int(c["num_classes"])).cpu().numpy()
    assert np.array_equal(stats, c["stats"][:, :, 1].astype(np.int64))
    # one call with every window equals the batched sequence (the order of a vote is (window, position) either way)
    once = E.SceneVotes(c["out_conf"].size)
    once.update(_cuda(c["conf"]), _cuda(c["pred"], torch.int32), _cuda(c["idx"], torch.int32), mapping)
    assert torch.equal(once.predictions, votes.predictions)


@pytest.mark.parametrize("name", ["shapenet_ties", "shapenet_plain", "shapenet_unvoted"])
def test_shape_merge_and_iou_equal_reference_numba_outputs(gold, name):
    from pvcnn_b200 import evaluate as E
    c = _case(gold, name)
    votes = E.SceneVotes(c["out_conf"].size)
    votes.update(_cuda(c["conf"])[None], _cuda(c["pred"], torch.int32)[None], _cuda(c["idx"], torch.int32)[None], None)
    assert np.array_equal(votes.predictions.cpu().numpy().astype(np.int64), c["out_pred"])
    assert np.array_equal(votes.confidences.cpu().numpy(), c["out_conf"])
    num_classes, c0, c1 = (int(v) for v in c["classes"])
    counts = votes.stats(c["gt"], num_classes, wrap_unvoted=False).cpu().numpy()
    assert abs(ev.shape_iou_from_counts(counts, c0, c1) - float(c["iou"])) < 1e-12       # shapenet eval.py:188-201
    assert abs(E.shape_iou(counts, c0, c1) - float(c["iou"])) < 1e-12


@pytest.mark.parametrize("n,num_classes", [(1, 13), (1000, 13), (1 << 20, 50), (3_000_001, 2048)])
def test_vote_stats_vs_oracle_both_conventions(n, num_classes):
    from pvcnn_b200 import _lib
    import ctypes
    g = np.random.default_rng(n % 1000 + num_classes)
    gt = g.integers(0, num_classes, size=n).astype(np.int32)
    pred = g.integers(-1, num_classes, size=n).astype(np.int32)
    agree = g.random(n) < 0.4
    pred[agree] = gt[agree]
    tg, tp = _cuda(gt), _cuda(pred)
    for wrap in (1, 0):
        out = torch.zeros((3, num_classes), dtype=torch.int64, device="cuda")
        _lib.call("pvcnn_vote_stats", ctypes.c_longlong(n), num_classes, wrap, tg, tp, out)
        want = ev.scene_counts(gt.astype(np.int64), pred.astype(np.int64), num_classes)
        if not wrap:
            want[1, num_classes - 1] -= int((pred == -1).sum())
        assert np.array_equal(out.cpu().numpy(), want)


class _PointwiseNet(torch.nn.Module):
    """a fixed per-point classifier standing in for the network (eval.py:173 only needs [B, classes, N] logits)"""

    def __init__(self, ch, num_classes):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.w = torch.nn.Parameter(torch.randn(num_classes, ch, generator=g))

    def forward(self, x):                                   # [B, ch, N] -> [B, classes, N], one fixed order per element
        return (self.w[None, :, :, None] * x[:, None, :, :]).sum(2)


@pytest.mark.parametrize("batch_size", [2, 5])
def test_evaluate_scene_file_equals_the_reference_loop(batch_size):
    """the whole loop of evaluate/s3dis/eval.py:139-182 on the device vs its restatement (oracle indices, literal numpy
    tiling, reference merge restatement pinned to numba) driven by the same per-vote confidences"""
    from pvcnn_b200 import evaluate as E
    g = np.random.default_rng(11)
    num_windows, p, ch, npo, num_votes, num_classes, scene_points = 5, 1500, 9, 512, 2, 13, 6000
    data = g.standard_normal((num_windows, p, ch)).astype(np.float32)
    npts = np.array([1500, 400, 1024, 37, 1200], np.int64)
    mapping = g.integers(0, scene_points, size=(num_windows, p)).astype(np.int64)
    gt = g.integers(0, num_classes, size=scene_points).astype(np.int64)
    net = _PointwiseNet(ch, num_classes).cuda().eval()
    votes = E.SceneVotes(scene_points)
    E.evaluate_scene_file(net, data, npts, mapping, votes, num_points=npo, num_votes=num_votes, batch_size=batch_size,
                          seed=99)
    extra = num_votes * -(-p // npo)
    nv = extra * npo
    conf_s = np.zeros(scene_points, np.float32)
    pred_s = np.full(scene_points, -1, np.int64)
    for lo in range(0, num_windows, batch_size):
        hi = min(lo + batch_size, num_windows)
        idx = ev.vote_indices(npts[lo:hi], nv, 99, lo)
        inputs = ev.vote_inputs(data[lo:hi], idx, npo)
        with torch.no_grad():
            conf, pred = E.softmax_max(net(_cuda(inputs)))
        ev.update_scene_predictions(conf.cpu().numpy().reshape(hi - lo, nv), pred.cpu().numpy().reshape(hi - lo, nv), idx,
                                    conf_s, pred_s, mapping, nv, hi - lo, lo)
    assert np.array_equal(votes.predictions.cpu().numpy().astype(np.int64), pred_s)
    assert np.array_equal(votes.confidences.cpu().numpy(), conf_s)
    assert np.array_equal(votes.stats(gt, num_classes).cpu().numpy(), ev.scene_counts(gt, pred_s, num_classes))


def test_evaluate_shape_equals_the_reference_loop():
    from pvcnn_b200 import evaluate as E
    g = np.random.default_rng(12)
    ch, n, npo, num_votes, num_classes, c0, c1 = 22, 2700, 1024, 3, 50, 12, 16
    ps = g.standard_normal((ch, n)).astype(np.float32)
    net = _PointwiseNet(ch, num_classes).cuda().eval()
    votes = E.evaluate_shape(net, ps, num_points=npo, num_votes=num_votes, start_class=c0, end_class=c1, seed=4,
                             shape_index=7)
    nv = num_votes * -(-n // npo) * npo
    idx = ev.vote_indices([n], nv, 4, 7)
    inputs = ev.shape_inputs(ps, idx[0], npo)
    with torch.no_grad():
        conf, pred = E.softmax_max(net(_cuda(inputs)), c0, c1)
    conf_s = np.zeros(n, np.float32)
    pred_s = np.full(n, -1, np.int64)
    ev.update_scene_predictions(conf.cpu().numpy().reshape(1, nv), pred.cpu().numpy().reshape(1, nv), idx, conf_s, pred_s,
                                None, nv, 1, 0)
    assert np.array_equal(votes.predictions.cpu().numpy().astype(np.int64), pred_s)
    assert (pred_s >= c0).all() and (pred_s < c1).all()       # nv >= n: every point is voted (shapenet eval.py:149-153)


def test_sample_windows_matches_choice_semantics_and_oracle():
    from pvcnn_b200 import evaluate as E
    g = np.random.default_rng(13)
    b, p, ch, k = 4, 6000, 9, 4096
    data = g.standard_normal((b, p, ch)).astype(np.float32)
    labels = g.integers(0, 13, size=(b, p)).astype(np.int64)
    npts = np.array([6000, 4096, 900, 5000], np.int64)
    out, lab = E.sample_windows(_cuda(data), _cuda(labels), npts, k, seed=21, first_window=3)
    idx = ev.window_indices(npts, k, 21, 3).astype(np.int64)
    assert out.shape == (b, ch, k) and lab.dtype == torch.int64
    for w in range(b):
        assert np.array_equal(out[w].cpu().numpy(), data[w][idx[w]].T)        # datasets/s3dis.py:88-89
        assert np.array_equal(lab[w].cpu().numpy(), labels[w][idx[w]])
        assert idx[w].max() < npts[w]
