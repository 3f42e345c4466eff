"""Test-time voting on the device: the host loops of the reference's evaluation scripts as kernels.

Mirrors evaluate/s3dis/eval.py:131-182 (`evaluate_scene_file`, `SceneVotes`) and evaluate/shapenet/eval.py:125-169
(`evaluate_shape`) and the window sampling of datasets/s3dis.py:86-89 (`sample_windows`).  The reference tiles,
shuffles and gathers the voted inputs with numpy per window, copies every batch's confidences / predictions back to the
host and merges them with numba loops; here the windows are uploaded once, every step is a kernel of
csrc/eval_voting.cu behind the C ABI (include/pvcnn_b200.h, "Test-time voting"), and only the final [3, classes]
counters are read back.  No CPU fallback: `_lib.call` raises without the CUDA library or with host tensors.

Random choices: the reference uses numpy's global generator (np.random.shuffle / np.random.choice); the device draws
the same distributions from a counter-based pseudo-random permutation keyed by (seed, window index), so a result does
not depend on how the windows are batched and is reproducible from the seed alone.
"""
import ctypes
import math

import torch

from . import _lib

__all__ = ["vote_indices", "window_indices", "vote_inputs", "softmax_max", "SceneVotes", "evaluate_scene_file",
           "evaluate_shape", "shape_iou", "sample_windows", "scenes_of_rank", "all_reduce_stats"]


_DEFAULT_DEVICE = "cuda"   # every kernel of this module needs a CUDA device; the tensors decide which one


def _default_device():
    return torch.device(_DEFAULT_DEVICE)


def _dev_i32(x, device):
    t = torch.as_tensor(x)
    return t.to(device=device, dtype=torch.int32).contiguous()


def vote_indices(num_points_in_window, total_num_voted_points, seed, first_window=0, device=None):
    """eval.py:161-164 for a batch of windows: int32 [b, nv] = tile(arange(n_w))[:nv], shuffled."""
    device = _default_device() if device is None else torch.device(device)
    n = _dev_i32(num_points_in_window, device).reshape(-1)
    out = torch.empty((n.numel(), int(total_num_voted_points)), dtype=torch.int32, device=n.device)
    _lib.call("pvcnn_vote_indices", n.numel(), int(total_num_voted_points), ctypes.c_ulonglong(seed & (2 ** 64 - 1)),
              int(first_window), n, out)
    return out


def window_indices(num_points_in_window, num_points, seed, first_window=0, device=None):
    """datasets/s3dis.py:86-87: np.random.choice(n_w, num_points, replace=(n_w < num_points)) -> int32 [b, num_points]"""
    device = _default_device() if device is None else torch.device(device)
    n = _dev_i32(num_points_in_window, device).reshape(-1)
    out = torch.empty((n.numel(), int(num_points)), dtype=torch.int32, device=n.device)
    _lib.call("pvcnn_window_indices", n.numel(), int(num_points), ctypes.c_ulonglong(seed & (2 ** 64 - 1)),
              int(first_window), n, out)
    return out


def vote_inputs(window_data, indices, num_points, channels_last=True, labels=None):
    """eval.py:166-171: window_data [b, P, ch] (channels_last, the h5 layout) or [b, ch, P] -> network input
    [b * extra, ch, num_points] with extra = nv / num_points; optionally labels [b, P] -> [b, nv]."""
    if window_data.dtype != torch.float32 or indices.dtype != torch.int32:
        raise RuntimeError("vote_inputs: window_data must be float32 and indices int32")
    window_data, indices = window_data.contiguous(), indices.contiguous()
    b, nv = indices.shape
    if channels_last:
        _, p, ch = window_data.shape
    else:
        _, ch, p = window_data.shape
    if window_data.shape[0] != b or nv % num_points:
        raise RuntimeError("vote_inputs: indices [b, nv] must match the windows and nv must be a multiple of num_points")
    extra = nv // num_points
    out = torch.empty((b * extra, ch, num_points), dtype=torch.float32, device=window_data.device)
    out_labels = None
    if labels is not None:
        labels = labels.to(torch.int32).contiguous()
        out_labels = torch.empty((b, nv), dtype=torch.int32, device=window_data.device)
    _lib.call("pvcnn_vote_gather", b, ch, p, extra, int(num_points), 1 if channels_last else 0, window_data, indices, out,
              labels, out_labels)
    return out if labels is None else (out, out_labels)


def softmax_max(logits, start_class=0, end_class=None):
    """F.softmax(logits, dim=1)[:, start:end].max(dim=1) (eval.py:173; shapenet eval.py:162-165, with the class offset
    already added): logits [b, c, n] -> (confidences fp32 [b, n], predictions int32 [b, n])."""
    if logits.dim() != 3 or logits.dtype != torch.float32:
        raise RuntimeError("softmax_max: logits must be float32 [b, c, n]")
    logits = logits.contiguous()
    b, c, n = logits.shape
    end_class = c if end_class is None else int(end_class)
    conf = torch.empty((b, n), dtype=torch.float32, device=logits.device)
    pred = torch.empty((b, n), dtype=torch.int32, device=logits.device)
    _lib.call("pvcnn_softmax_max", b, c, n, int(start_class), end_class, logits, conf, pred)
    return conf, pred


class SceneVotes:
    """The per-scene state of eval.py:136-137 (`confidences` zeros, `predictions` -1) on the device, and the merge of
    `update_scene_predictions` (:189-204).  One 64-bit word per scene point holds (confidence bits, ~sequence number of
    the vote), so an atomic max realises "replace iff strictly more confident, earliest vote wins ties"."""

    def __init__(self, total_num_points, device=None):
        device = _default_device() if device is None else torch.device(device)
        self.total_num_points = int(total_num_points)
        self._keys = torch.empty(self.total_num_points, dtype=torch.int64, device=device)   # raw uint64 words
        self._pred = torch.empty(self.total_num_points, dtype=torch.int32, device=device)
        self._votes = 0
        _lib.call("pvcnn_vote_reset", ctypes.c_longlong(self.total_num_points), self._keys, self._pred)

    def update(self, confidences, predictions, indices, window_to_scene_mapping=None):
        """confidences fp32 / predictions int32 / indices int32, all [b, nv]; window_to_scene_mapping int [b, P] = the rows
        of `indices_split_to_full` of these windows (None: shapenet, the index is the point)."""
        b, nv = indices.shape
        confidences = confidences.reshape(b, nv).to(torch.float32).contiguous()
        predictions = predictions.reshape(b, nv).to(torch.int32).contiguous()
        indices = indices.to(torch.int32).contiguous()
        p = 1
        if window_to_scene_mapping is not None:
            window_to_scene_mapping = window_to_scene_mapping.to(device=indices.device, dtype=torch.int32).contiguous()
            if window_to_scene_mapping.shape[0] != b:
                raise RuntimeError("SceneVotes.update: one mapping row per window")
            p = window_to_scene_mapping.shape[1]
        if self._votes + b * nv > 2 ** 32:
            raise RuntimeError("SceneVotes: more than 2^32 votes into one scene")
        _lib.call("pvcnn_vote_merge", b, nv, p, ctypes.c_longlong(self.total_num_points), ctypes.c_uint(self._votes),
                  confidences, predictions, indices, window_to_scene_mapping, self._keys, self._pred)
        self._votes += b * nv

    @property
    def predictions(self):
        """int32 [total_num_points]; -1 where no vote with a positive confidence arrived (eval.py:137)"""
        return self._pred

    @property
    def confidences(self):
        out = torch.empty(self.total_num_points, dtype=torch.float32, device=self._keys.device)
        _lib.call("pvcnn_vote_confidences", ctypes.c_longlong(self.total_num_points), self._keys, out)
        return out

    def stats(self, ground_truth, num_classes, wrap_unvoted=True):
        """update_stats (eval.py:207-215) for this scene: int64 [3, num_classes] = (ground-truth count, prediction count,
        agreement).  wrap_unvoted=True keeps the reference's numba behaviour of counting unvoted points in the last class
        of row 1; False counts them nowhere (what shapenet's per-class IoU needs)."""
        gt = _dev_i32(ground_truth, self._pred.device).reshape(-1)
        if gt.numel() != self.total_num_points:
            raise RuntimeError("SceneVotes.stats: one ground-truth label per scene point")
        out = torch.zeros((3, int(num_classes)), dtype=torch.int64, device=self._pred.device)
        _lib.call("pvcnn_vote_stats", ctypes.c_longlong(self.total_num_points), int(num_classes), 1 if wrap_unvoted else 0,
                  gt, self._pred, out)
        return out


def evaluate_scene_file(model, scene_data, scene_num_points, window_to_scene_mapping, votes, *, num_points, num_votes,
                        batch_size, seed=0, first_window=0):
    """One h5 file of a scene, as eval.py:139-179: scene_data [num_windows, P, ch] fp32, scene_num_points [num_windows],
    window_to_scene_mapping [num_windows, P] -> merged into `votes` (a SceneVotes).  The windows are uploaded once; per
    batch of windows: indices -> gathered inputs -> model -> softmax-max -> merge, all on the device."""
    device = votes.predictions.device
    scene_data = torch.as_tensor(scene_data).to(device=device, dtype=torch.float32).contiguous()
    num_windows, max_points, _ = scene_data.shape
    npts = _dev_i32(scene_num_points, device).reshape(-1)
    mapping = _dev_i32(window_to_scene_mapping, device)
    extra_batch_size = num_votes * math.ceil(max_points / num_points)                      # eval.py:146
    total_num_voted_points = extra_batch_size * num_points                                 # eval.py:147
    for lo in range(0, num_windows, batch_size):                                           # eval.py:149
        hi = min(lo + batch_size, num_windows)
        idx = vote_indices(npts[lo:hi], total_num_voted_points, seed, first_window + lo, device)
        inputs = vote_inputs(scene_data[lo:hi], idx, num_points)
        with torch.no_grad():
            conf, pred = softmax_max(model(inputs))                                        # eval.py:172-173
        votes.update(conf.view(hi - lo, total_num_voted_points), pred.view(hi - lo, total_num_voted_points), idx,
                     mapping[lo:hi])
    return votes


def evaluate_shape(model, point_set, *, num_points, num_votes, start_class, end_class, seed=0, shape_index=0):
    """One ShapeNet shape, as shapenet eval.py:149-168: point_set [ch, n] fp32 -> SceneVotes over its n points."""
    point_set = torch.as_tensor(point_set)
    device = point_set.device if point_set.is_cuda else _default_device()
    point_set = point_set.to(device=device, dtype=torch.float32).contiguous()
    n = point_set.shape[1]
    extra_batch_size = num_votes * math.ceil(n / num_points)                               # shapenet eval.py:149
    total_num_voted_points = extra_batch_size * num_points
    votes = SceneVotes(n, device)
    idx = vote_indices([n], total_num_voted_points, seed, shape_index, device)
    inputs = vote_inputs(point_set[None], idx, num_points, channels_last=False)
    with torch.no_grad():
        conf, pred = softmax_max(model(inputs), start_class, end_class)                    # shapenet eval.py:161-165
    votes.update(conf.view(1, -1), pred.view(1, -1), idx, None)
    return votes


def shape_iou(counts, start_class, end_class):
    """update_stats of evaluate/shapenet/eval.py:188-201 from the [3, classes] counters of `SceneVotes.stats(...,
    wrap_unvoted=False)`: mean over the shape's part classes of intersection / union, 1 where the union is empty
    (union = |gt == i| + |pred == i| - |both|).  Host arithmetic on <= 50 numbers."""
    c = torch.as_tensor(counts).to("cpu", torch.float64)[:, int(start_class):int(end_class)]
    union = c[0] + c[1] - c[2]
    iou = torch.where(union == 0, torch.ones_like(union), c[2] / union.clamp(min=1))
    return float(iou.mean())


def sample_windows(window_data, window_labels, window_num_points, num_points, seed=0, first_window=0):
    """datasets/s3dis.py:86-94 for a batch of windows already on the device: window_data [b, P, ch], window_labels [b, P],
    window_num_points [b] -> (data [b, ch, num_points] fp32, labels [b, num_points] int64)."""
    idx = window_indices(window_num_points, num_points, seed, first_window, window_data.device)
    data, labels = vote_inputs(window_data, idx, num_points, labels=window_labels)
    return data, labels.long()


# ---- multi-GPU: scenes (shapes) are independent units ------------------------------------------------
def scenes_of_rank(num_scenes, rank=None, world=None):
    """Scene indices evaluated by this rank: round-robin over the ranks (scene sizes vary by an order of magnitude, so
    interleaving balances better than contiguous blocks).  No data-path collective: every rank runs whole scenes."""
    import torch.distributed as dist
    if rank is None or world is None:
        on = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank() if on else 0
        world = dist.get_world_size() if on else 1
    return list(range(rank, int(num_scenes), world))


def all_reduce_stats(stats, group=None):
    """The one exchange step of a sharded evaluation: `stats` [3, classes, scenes] (eval.py:131) holds this rank's scenes'
    columns and zeros elsewhere; a SUM all-reduce gives every rank the full table.  int64 counters, NCCL or gloo."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
