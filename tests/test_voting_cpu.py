"""CPU tests of the test-time voting oracle (oracle/eval_voting.py).

1. Pinned: the merge / statistics restatements equal the REFERENCE's own numba functions
   (evaluate/s3dis/eval.py:189-215, evaluate/shapenet/eval.py:177-201) on tests/golden/ref_voting_golden.npz, which
   tests/golden/make_voting_golden.py produced by importing and running the unmodified reference files.
2. Generator: the counter-based permutation is a permutation; the voted indices have the multiset the reference's
   tile + shuffle produces; window sampling is a subset without replacement / in range with replacement.
3. The literal input tiling of eval.py:157-171 equals the index formula the device kernel uses.
"""
import os

import numpy as np
import pytest

from oracle import eval_voting as ev

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_voting_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(G)


def _case(gold, name):
    return {k.split(".", 1)[1]: gold[k] for k in gold.files if k.startswith(name + ".")}


@pytest.mark.parametrize("name", ["s3dis_ties", "s3dis_dense", "s3dis_sparse"])
def test_scene_merge_and_stats_equal_reference_numba(gold, name):
    c = _case(gold, name)
    num_windows, nv = c["conf"].shape
    bs = int(c["batch_size"])
    scene_points = c["out_conf"].size
    conf = np.zeros(scene_points, np.float32)
    pred = np.full(scene_points, -1, np.int64)
    for lo in range(0, num_windows, bs):
        hi = min(lo + bs, num_windows)
        ev.update_scene_predictions(c["conf"][lo:hi], c["pred"][lo:hi], c["idx"][lo:hi], conf, pred, c["mapping"], nv,
                                    hi - lo, lo)
    assert np.array_equal(conf, c["out_conf"])
    assert np.array_equal(pred, c["out_pred"])
    stats = np.zeros_like(c["stats"])
    ev.update_stats(stats, c["gt"], pred, 1, scene_points)
    assert np.array_equal(stats, c["stats"])
    if name == "s3dis_sparse":   # the quirk is exercised: unvoted points sit in the last class of row 1
        assert (pred == -1).sum() > 0
        assert stats[1, -1, 1] >= (pred == -1).sum()
    assert np.array_equal(ev.scene_counts(c["gt"], pred, int(c["num_classes"])), c["stats"][:, :, 1].astype(np.int64))


@pytest.mark.parametrize("name", ["shapenet_ties", "shapenet_plain", "shapenet_unvoted"])
def test_shape_merge_and_iou_equal_reference_numba(gold, name):
    c = _case(gold, name)
    n = c["out_conf"].size
    conf = np.zeros(n, np.float32)
    pred = np.full(n, -1, np.int64)
    ev.update_scene_predictions(c["conf"][None], c["pred"][None], c["idx"][None], conf, pred, None, c["conf"].size, 1, 0)
    assert np.array_equal(conf, c["out_conf"])
    assert np.array_equal(pred, c["out_pred"])
    num_classes, c0, c1 = (int(v) for v in c["classes"])
    assert abs(ev.shape_iou(c["gt"], pred, c0, c1) - float(c["iou"])) < 1e-12
    # the same IoU from the [3, classes] counters the device produces (unvoted points counted nowhere)
    counts = np.zeros((3, num_classes), np.int64)
    np.add.at(counts[0], c["gt"], 1)
    np.add.at(counts[1], pred[pred >= 0], 1)
    np.add.at(counts[2], c["gt"][c["gt"] == pred], 1)
    assert abs(ev.shape_iou_from_counts(counts, c0, c1) - float(c["iou"])) < 1e-12
    from pvcnn_b200 import evaluate                       # the product's host-side reduction of the same counters
    assert abs(evaluate.shape_iou(counts, c0, c1) - float(c["iou"])) < 1e-12


def test_sequential_rule_first_vote_wins_ties():
    """hand case of eval.py:201: strictly larger replaces, so the earliest of equal confidences stays"""
    conf = np.zeros(3, np.float32)
    pred = np.full(3, -1, np.int64)
    bc = np.array([[0.5, 0.5, 0.25, 0.0]], np.float32)
    bp = np.array([[7, 9, 3, 4]], np.int64)
    idx = np.array([[0, 0, 1, 2]], np.int64)
    ev.update_scene_predictions(bc, bp, idx, conf, pred, None, 4, 1, 0)
    assert pred.tolist() == [7, 3, -1] and conf.tolist() == [0.5, 0.25, 0.0]
    ev.update_scene_predictions(np.array([[0.5, 0.75]], np.float32), np.array([[1, 2]], np.int64),
                                np.array([[0, 1]], np.int64), conf, pred, None, 2, 1, 0)
    assert pred.tolist() == [7, 2, -1]     # an equal confidence in a later call does not replace


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 16, 17, 255, 256, 257, 1000, 4096, 65536, 65537, 100003])
def test_feistel_perm_is_a_permutation(n):
    p = ev.feistel_perm(n, 1234567, 3)
    assert p.min() == 0 and p.max() == n - 1
    assert np.unique(p).size == n
    if n >= 256:   # not the identity, differs between streams and seeds
        assert (p != np.arange(n)).mean() > 0.9
        assert (p != ev.feistel_perm(n, 1234567, 4)).mean() > 0.9
        assert (p != ev.feistel_perm(n, 1234568, 3)).mean() > 0.9
    # evaluating single positions equals the full table (what each device thread does)
    xs = np.array([0, n // 2, n - 1])
    assert np.array_equal(ev.feistel_perm(n, 1234567, 3, x=xs), p[xs])


def test_feistel_perm_looks_uniform():
    """position of element 0 over many streams is spread over the whole range (chi-square over 16 bins)"""
    n = 1000
    pos = np.array([int(ev.feistel_perm(n, 99, s, x=[0])[0]) for s in range(4000)])
    hist = np.bincount(pos * 16 // n, minlength=16)
    chi2 = ((hist - 250.0) ** 2 / 250.0).sum()
    assert chi2 < 45.0   # 15 degrees of freedom: P(chi2 > 45) < 1e-4


def test_vote_indices_multiset_matches_tile_and_shuffle():
    num_points = np.array([700, 4096, 1, 5000, 0])
    nv = 3 * 4096
    idx = ev.vote_indices(num_points, nv, 42, first_window=10)
    for w, n in enumerate(num_points):
        if n == 0:
            assert not idx[w].any()
            continue
        counts = np.bincount(idx[w], minlength=n)
        expect = np.bincount(np.tile(np.arange(n), -(-nv // n))[:nv], minlength=n)   # eval.py:161-163 (num_repeats, np.tile, [:nv])
        assert np.array_equal(np.sort(counts), np.sort(expect))
        assert counts.size == n and counts.min() >= nv // n and counts.max() <= -(-nv // n)
    # batching invariance: window 12 alone equals row 2 of the batch that started at window 10
    assert np.array_equal(ev.vote_indices(num_points[2:4], nv, 42, first_window=12), idx[2:4])


def test_window_indices_choice_semantics():
    num_points = np.array([5000, 4096, 100, 1, 0])
    k = 4096
    idx = ev.window_indices(num_points, k, 7)
    assert np.unique(idx[0]).size == k and idx[0].max() < 5000           # without replacement
    assert np.array_equal(np.sort(idx[1]), np.arange(4096))              # n == k: a permutation
    assert idx[2].max() < 100 and np.unique(idx[2]).size > 90            # with replacement, covers the window
    assert not idx[3].any() and not idx[4].any()


def test_vote_inputs_literal_tiling_equals_index_formula():
    g = np.random.default_rng(5)
    b, p, ch, npo, extra = 3, 50, 9, 16, 4
    data = g.standard_normal((b, p, ch)).astype(np.float32)
    idx = ev.vote_indices(np.array([50, 20, 33]), extra * npo, 11)
    out = ev.vote_inputs(data, idx, npo)
    assert out.shape == (b * extra, ch, npo)
    for w in range(b):
        for e in range(extra):
            for j in (0, 7, 15):
                assert np.array_equal(out[w * extra + e, :, j], data[w, idx[w, e * npo + j]])
    ps = g.standard_normal((ch, p)).astype(np.float32)
    so = ev.shape_inputs(ps, idx[0], npo)
    assert so.shape == (extra, ch, npo)
    assert np.array_equal(so[2, :, 5], ps[:, idx[0, 2 * npo + 5]])


def test_kernel_source_under_cpu_emulation():
    """tests/tools/emulate_voting_on_cpu.py compiles the UNMODIFIED csrc/eval_voting.cu with g++ behind a small CUDA shim
    (thread-local blockIdx / threadIdx, atomics as GCC builtins, real threads where the kernel synchronises) and runs the
    whole GPU test file against it: kernel logic, launch geometry and the ctypes argument order are checked without a GPU.
    Test infrastructure only -- the product library has no CPU path (tests/test_abi_cpu.py::test_voting_host_api_has_no_cpu_path)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tests", "tools", "emulate_voting_on_cpu.py")],
                       capture_output=True, text=True, cwd=root, timeout=900,
                       env=dict(os.environ, PVCNN_TEST_BUDGET_S="0"))
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-1500:])
    assert "30 passed" in p.stdout
