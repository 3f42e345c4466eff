"""CPU oracle of the test-time voting path -- TEST INFRASTRUCTURE, never the product path.

Restates, in numpy, what the reference's evaluation loops compute on the host:
  evaluate/s3dis/eval.py:149-215      (tile / shuffle / gather, softmax-max, update_scene_predictions, update_stats)
  evaluate/shapenet/eval.py:149-201   (same scheme per shape, class range of the shape, per-shape IoU)
  datasets/s3dis.py:86-89             (np.random.choice window sampling)
The merge / statistics functions are PINNED against the reference's own numba functions: tests/golden/
make_voting_golden.py imports them from /root/reference, runs them on seeded inputs (ties, unvoted points, several
batches) and the outputs are committed as tests/golden/ref_voting_golden.npz (tests/test_voting_cpu.py).
The random choices of the reference come from numpy's global generator, which a device cannot reproduce; the product
draws them from a counter-based pseudo-random permutation instead (pvcnn_b200/csrc/eval_voting.cu), restated here bit
for bit (`feistel_perm`), and the tests check the distribution-free properties of the reference's scheme on top
(every index of a window appears floor or ceil(nv / n) times; subsets without replacement are distinct).
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
ROUNDS = 6


def _mix(seed, stream, j):
    """splitmix64 finaliser of (seed, stream, index): eval_voting.cu::vt_mix"""
    with np.errstate(over="ignore"):
        x = (np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ (np.asarray(stream, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
             ^ (np.asarray(j, dtype=np.uint64) * np.uint64(0x94D049BB133111EB)))
        x = x ^ (x >> np.uint64(30)); x = x * np.uint64(0xBF58476D1CE4E5B9)
        x = x ^ (x >> np.uint64(27)); x = x * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def _round(r, key):
    """eval_voting.cu::vt_round on uint32 lanes (murmur3 finaliser of r * golden + key)"""
    m32 = np.uint64(0xFFFFFFFF)
    x = (r.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(key)) & m32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x85EBCA6B)) & m32
    x ^= x >> np.uint64(13); x = (x * np.uint64(0xC2B2AE35)) & m32
    x ^= x >> np.uint64(16)
    return x


def feistel_perm(n, seed, stream, x=None):
    """perm(x) for x in [0, n) (default: all of them): six-round balanced Feistel network over 2h bits, cycle-walked
    back into [0, n) -- eval_voting.cu::vt_perm."""
    n = int(n)
    assert 0 < n <= 2 ** 31
    h = 1
    while (1 << (2 * h)) < n:
        h += 1
    mask = np.uint64((1 << h) - 1)
    keys = [int(_mix(seed, stream, r) >> np.uint64(32)) for r in range(ROUNDS)]
    x = (np.arange(n, dtype=np.uint64) if x is None else np.asarray(x, dtype=np.uint64)).copy()
    todo = np.ones(x.shape, bool)
    while todo.any():
        v = x[todo]
        l, r = v >> np.uint64(h), v & mask
        for k in range(ROUNDS):
            l, r = r, l ^ (_round(r, keys[k]) & mask)
        v = (l << np.uint64(h)) | r
        x[todo] = v
        todo[todo] = v >= np.uint64(n)
    return x.astype(np.int64)


def vote_indices(num_points, nv, seed, first_window=0):
    """evaluate/s3dis/eval.py:161-164 with the device generator: tile(arange(n_w))[:nv] shuffled = perm_nv(p) mod n_w.
    num_points [b] -> int32 [b, nv]"""
    num_points = np.asarray(num_points).reshape(-1)
    out = np.zeros((num_points.size, nv), np.int32)
    for w, n in enumerate(num_points):
        if n > 0:
            out[w] = feistel_perm(nv, seed, first_window + w) % int(n)
    return out


def window_indices(num_points, k, seed, first_window=0):
    """datasets/s3dis.py:86-87 with the device generator: a uniform k-subset in random order when n_w >= k, k independent
    uniform draws otherwise.  num_points [b] -> int32 [b, k]"""
    num_points = np.asarray(num_points).reshape(-1)
    out = np.zeros((num_points.size, k), np.int32)
    j = np.arange(k, dtype=np.uint64)
    for w, n in enumerate(num_points):
        n = int(n)
        if n >= k:
            out[w] = feistel_perm(n, seed, first_window + w, x=j)
        elif n > 0:
            u = _mix(seed ^ 0xD1B54A32D192ED03, first_window + w, j) >> np.uint64(32)
            out[w] = ((u * np.uint64(n)) >> np.uint64(32)).astype(np.int32)
    return out


def vote_inputs(window_data, indices, num_points):
    """evaluate/s3dis/eval.py:157-171, literally: window_data [b, P, ch] -> [b * extra, ch, num_points]"""
    b, nv = indices.shape
    ch = window_data.shape[-1]
    batched = np.zeros((b, nv, ch), np.float32)
    for w in range(b):
        batched[w] = window_data[w][indices[w]]
    extra = nv // num_points
    return np.ascontiguousarray(batched.reshape((b * extra, num_points, -1)).transpose(0, 2, 1))


def shape_inputs(point_set, indices, num_points):
    """evaluate/shapenet/eval.py:158-160, literally: point_set [ch, P] -> [extra, ch, num_points]"""
    extra = indices.size // num_points
    return np.ascontiguousarray(point_set[:, indices].reshape(-1, extra, num_points).transpose(1, 0, 2))


def softmax_max(logits, c0=0, c1=None):
    """evaluate/s3dis/eval.py:173 / shapenet eval.py:162-165 through the reference's own torch calls (CPU fp32)"""
    import torch
    import torch.nn.functional as F
    t = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32))
    c1 = t.shape[1] if c1 is None else c1
    conf, pred = F.softmax(t, dim=1)[:, c0:c1, :].max(dim=1)
    return conf.numpy(), (pred + c0).numpy().astype(np.int32)


def update_scene_predictions(batched_confidences, batched_predictions, batched_shuffled_point_indices,
                             scene_confidences, scene_predictions, window_to_scene_mapping, total_num_voted_points,
                             batch_size, min_window_index):
    """evaluate/s3dis/eval.py:189-204, same signature, in place.  The sequential rule (replace iff strictly larger) means:
    per scene point the winner of a call is its most confident vote, the earliest one among equals, and it replaces the
    stored entry iff it is strictly more confident.  window_to_scene_mapping=None: shapenet's update_shape_predictions
    (shapenet eval.py:177-185, the shuffled index is the point)."""
    conf = np.asarray(batched_confidences).reshape(batch_size, total_num_voted_points)
    pred = np.asarray(batched_predictions).reshape(batch_size, total_num_voted_points)
    idx = np.asarray(batched_shuffled_point_indices).reshape(batch_size, total_num_voted_points)
    if window_to_scene_mapping is None:
        pts = idx.reshape(-1).astype(np.int64)
    else:
        rows = np.asarray(window_to_scene_mapping)[min_window_index:min_window_index + batch_size]
        pts = np.take_along_axis(rows, idx.astype(np.int64), axis=1).reshape(-1).astype(np.int64)
    conf, pred = conf.reshape(-1), pred.reshape(-1)
    order = np.arange(conf.size)
    keep = ~np.isnan(conf)
    pts, conf, pred, order = pts[keep], conf[keep], pred[keep], order[keep]
    srt = np.lexsort((order, -conf.astype(np.float64), pts))    # by point, then confidence descending, then order
    pts, conf, pred = pts[srt], conf[srt], pred[srt]
    first = np.ones(pts.size, bool)
    first[1:] = pts[1:] != pts[:-1]
    pts, conf, pred = pts[first], conf[first], pred[first]
    win = conf > scene_confidences[pts]
    scene_confidences[pts[win]] = conf[win]
    scene_predictions[pts[win]] = pred[win]


def update_stats(stats, ground_truth, predictions, scene_index, total_num_points_in_scene):
    """evaluate/s3dis/eval.py:207-215, in place; a prediction of -1 indexes the last class (numba wraps negative indices)"""
    gt = np.asarray(ground_truth[:total_num_points_in_scene]).astype(np.int64)
    pd = np.asarray(predictions[:total_num_points_in_scene]).astype(np.int64)
    nc = stats.shape[1]
    np.add.at(stats[0, :, scene_index], gt, 1)
    np.add.at(stats[1, :, scene_index], np.where(pd < 0, pd + nc, pd), 1)
    np.add.at(stats[2, :, scene_index], gt[gt == pd], 1)


def scene_counts(ground_truth, predictions, num_classes):
    """the [3, num_classes] counters of one scene (update_stats on a zeroed column)"""
    stats = np.zeros((3, num_classes, 1))
    update_stats(stats, ground_truth, predictions, 0, len(ground_truth))
    return stats[:, :, 0].astype(np.int64)


def shape_iou(ground_truth, predictions, start_class, end_class):
    """evaluate/shapenet/eval.py:188-201: mean over the shape's part classes of intersection / union (1 when both empty)"""
    iou = 0.0
    for i in range(start_class, end_class):
        igt, ipd = ground_truth == i, predictions == i
        union = np.sum(igt | ipd)
        iou += 1 if union == 0 else np.sum(igt & ipd) / union
    return iou / (end_class - start_class)


def shape_iou_from_counts(counts, start_class, end_class):
    """the same IoU from the [3, classes] counters (union = |gt| + |pred| - |both|)"""
    iou = 0.0
    for i in range(start_class, end_class):
        union = counts[0, i] + counts[1, i] - counts[2, i]
        iou += 1 if union == 0 else counts[2, i] / union
    return iou / (end_class - start_class)
