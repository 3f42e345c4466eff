"""Generates tests/golden/ref_voting_golden.npz from the UNMODIFIED reference evaluation code.

Run in the build container (needs /root/reference and numba; no GPU): the merge / statistics functions of
evaluate/s3dis/eval.py (`update_scene_predictions` :189-204, `update_stats` :207-215) and evaluate/shapenet/eval.py
(`update_shape_predictions` :177-185, `update_stats` :188-201) are imported from the reference files where they lie and
executed (numba-compiled, as the reference runs them) on seeded inputs that exercise what a parallel merge can get
wrong: heavily tied confidences, several sequential batches into one scene, scene points that never receive a vote
(prediction stays -1 and is counted in the last class by numba's wrap-around), confidences equal to the initial 0.

    python tests/golden/make_voting_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def s3dis_case(ref, g, scene_points, num_windows, window_points, nv, batch_size, num_classes, quant):
    """one scene merged batch by batch exactly like evaluate/s3dis/eval.py:149-182"""
    mapping = g.integers(0, scene_points, size=(num_windows, window_points)).astype(np.int64)
    num_pts = g.integers(max(1, window_points // 3), window_points + 1, size=num_windows).astype(np.int64)
    confidences = np.zeros(scene_points, np.float32)
    predictions = np.full(scene_points, -1, np.int64)
    conf_all = np.zeros((num_windows, nv), np.float32)
    pred_all = np.zeros((num_windows, nv), np.int64)
    idx_all = np.zeros((num_windows, nv), np.int64)
    for lo in range(0, num_windows, batch_size):
        hi = min(lo + batch_size, num_windows)
        bs = hi - lo
        conf = g.random((bs, nv), dtype=np.float32)
        if quant:
            conf = np.floor(conf * quant).astype(np.float32) / np.float32(quant)   # ties, and exact zeros
        pred = g.integers(0, num_classes, size=(bs, nv)).astype(np.int64)
        idx = np.stack([g.integers(0, num_pts[lo + w], size=nv) for w in range(bs)]).astype(np.int64)
        ref.update_scene_predictions(conf, pred, idx, confidences, predictions, mapping, nv, bs, lo)
        conf_all[lo:hi], pred_all[lo:hi], idx_all[lo:hi] = conf, pred, idx
    gt = g.integers(0, num_classes, size=scene_points).astype(np.int64)
    stats = np.zeros((3, num_classes, 2))
    ref.update_stats(stats, gt, predictions, 1, scene_points)
    return dict(mapping=mapping, conf=conf_all, pred=pred_all, idx=idx_all, batch_size=np.int64(batch_size),
                out_conf=confidences, out_pred=predictions, gt=gt, stats=stats, num_classes=np.int64(num_classes))


def shapenet_case(ref, g, n, nv, num_classes, start_class, end_class, quant):
    """one shape as in evaluate/shapenet/eval.py:149-169"""
    confidences = np.zeros(n, np.float32)
    predictions = np.full(n, -1, np.int64)
    conf = g.random(nv, dtype=np.float32)
    if quant:
        conf = np.floor(conf * quant).astype(np.float32) / np.float32(quant)
    pred = g.integers(start_class, end_class, size=nv).astype(np.int64)
    idx = g.integers(0, n, size=nv).astype(np.int64)
    ref.update_shape_predictions(conf, pred, idx, confidences, predictions, nv)
    gt = g.integers(start_class, end_class, size=n).astype(np.int64)
    stats = np.zeros((4, 2))
    ref.update_stats(stats, gt, predictions, 2, start_class, end_class)
    return dict(conf=conf, pred=pred, idx=idx, out_conf=confidences, out_pred=predictions, gt=gt,
                iou=np.float64(stats[2, 0]), classes=np.array([num_classes, start_class, end_class], np.int64))


def main():
    sys.path.insert(0, REF)
    s3 = _load("ref_s3dis_eval", "evaluate/s3dis/eval.py")
    sn = _load("ref_shapenet_eval", "evaluate/shapenet/eval.py")
    g = np.random.default_rng(1588147245)
    out = {}
    cases = {
        "s3dis_ties": s3dis_case(s3, g, 3000, 7, 600, 1024, 3, 13, 8),
        "s3dis_dense": s3dis_case(s3, g, 500, 5, 400, 2048, 2, 13, 0),
        "s3dis_sparse": s3dis_case(s3, g, 20000, 4, 300, 512, 4, 13, 4),       # most scene points never voted
        "shapenet_ties": shapenet_case(sn, g, 2500, 4096, 50, 12, 16, 16),
        "shapenet_plain": shapenet_case(sn, g, 700, 2048, 50, 0, 4, 0),
        "shapenet_unvoted": shapenet_case(sn, g, 5000, 1024, 50, 47, 50, 8),
    }
    for name, d in cases.items():
        for k, v in d.items():
            out["%s.%s" % (name, k)] = v
    path = os.path.join(HERE, "ref_voting_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
