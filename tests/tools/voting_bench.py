"""Times the test-time voting steps AROUND the network (evaluate/s3dis/eval.py:149-179 without the model call of :173): host arm vs device arm.

Workload = one batch of the reference's S3DIS evaluation: batch_size 10 windows (configs/s3dis/__init__.py:22) of up to
8192 points x 9 channels (data/s3dis/prepare_data.py:88), num_points 4096, num_votes 1 -> extra_batch_size 2, 8192 voted
points per window, 13 classes; scene of 1 M points.

  host arm    the reference's own steps on the host cores: np.tile / np.random.shuffle / fancy indexing per window
              (:155-171) and the merge through the reference's numba function when /root/reference is importable
              (else the oracle restatement) -- what the reference pays on the HOST per batch next to the network, not
              counting its H2D / D2H copies.  (softmax + max run on the device in the reference too, :173; the torch-CPU
              time of that step is printed for information and is not part of `ms_total`.)
  device arm  pvcnn_b200.evaluate: vote_indices + vote_inputs + softmax_max + SceneVotes.update, CUDA events.

  python tests/tools/voting_bench.py [--host-only] [--iters 20]
"""
import argparse
import importlib.util
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

BATCH, P, CH, NPO, VOTES, CLASSES, SCENE = 10, 8192, 9, 4096, 1, 13, 1_000_000


def workload(g):
    data = g.standard_normal((BATCH, P, CH)).astype(np.float32)
    npts = g.integers(P // 4, P + 1, size=BATCH).astype(np.int64)
    mapping = g.integers(0, SCENE, size=(BATCH, P)).astype(np.int64)
    extra = VOTES * math.ceil(P / NPO)
    logits = g.standard_normal((BATCH * extra, CLASSES, NPO)).astype(np.float32)
    return data, npts, mapping, extra, logits


def host_arm(iters):
    import torch
    import torch.nn.functional as F
    merge, kind = None, "oracle restatement"
    path = "/root/reference/evaluate/s3dis/eval.py"
    if os.path.exists(path):
        sys.path.insert(0, "/root/reference")
        spec = importlib.util.spec_from_file_location("ref_s3dis_eval", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        merge, kind = mod.update_scene_predictions, "reference numba function"
    else:
        sys.path.insert(0, os.path.join(ROOT))
        from oracle import eval_voting
        merge = eval_voting.update_scene_predictions
    g = np.random.default_rng(0)
    data, npts, mapping, extra, logits = workload(g)
    nv = extra * NPO
    conf_s = np.zeros(SCENE, np.float32)
    pred_s = np.full(SCENE, -1, np.int64)
    t_tile = t_soft = t_merge = 0.0
    for it in range(iters + 1):
        t0 = time.perf_counter()
        batched_inputs = np.zeros((BATCH, nv, CH), dtype=np.float32)                       # eval.py:157-166
        batched_idx = np.zeros((BATCH, nv), dtype=np.int64)
        for w in range(BATCH):
            n = npts[w]
            idx = np.tile(np.arange(n), math.ceil(nv / n))[:nv]
            np.random.shuffle(idx)
            batched_idx[w] = idx
            batched_inputs[w] = data[w][idx]
        inputs = torch.from_numpy(batched_inputs.reshape((BATCH * extra, NPO, -1)).transpose(0, 2, 1)).float().contiguous()
        t1 = time.perf_counter()
        conf, pred = F.softmax(torch.from_numpy(logits), dim=1).max(dim=1)                 # eval.py:173-175
        conf, pred = conf.view(BATCH, nv).numpy(), pred.view(BATCH, nv).numpy()
        t2 = time.perf_counter()
        merge(conf, pred, batched_idx, conf_s, pred_s, mapping, nv, BATCH, 0)              # eval.py:177-179
        t3 = time.perf_counter()
        if it:   # first pass = numba compilation / warm-up
            t_tile += t1 - t0
            t_soft += t2 - t1
            t_merge += t3 - t2
        del inputs
    # the numpy restatement of the merge is a checker (a lexsort), not a stand-in for the reference's compiled numba loop:
    # where the reference tree is absent (the GPU box) the total counts the tiling only (97 % of the reference's host time)
    counted = (t_tile + t_merge) if kind == "reference numba function" else t_tile
    return {"arm": "host", "merge": kind, "merge_counted_in_total": kind == "reference numba function",
            "threads_torch": torch.get_num_threads(), "cores": os.cpu_count(),
            "ms_tile_shuffle_gather": t_tile / iters * 1e3, "ms_softmax_max_torch_cpu_info_only": t_soft / iters * 1e3,
            "ms_merge": t_merge / iters * 1e3, "ms_total": counted / iters * 1e3,
            "voted_points_per_batch": BATCH * nv}


def device_arm(iters):
    import torch
    from pvcnn_b200 import evaluate as E
    g = np.random.default_rng(0)
    data, npts, mapping, extra, logits = workload(g)
    nv = extra * NPO
    data_d = torch.from_numpy(data).cuda()
    map_d = torch.from_numpy(mapping).cuda().int()
    npts_d = torch.from_numpy(npts).cuda().int()
    logits_d = torch.from_numpy(logits).cuda()
    votes = E.SceneVotes(SCENE)

    def once(i):
        idx = E.vote_indices(npts_d, nv, 7, i * BATCH)
        E.vote_inputs(data_d, idx, NPO)
        conf, pred = E.softmax_max(logits_d)
        votes.update(conf.view(BATCH, nv), pred.view(BATCH, nv), idx, map_d)

    for i in range(3):
        once(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        once(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # algorithmic bytes per batch: indices 4, gather 4 + 8*CH, softmax 4*CLASSES + 8, merge 16 + 8 (atomic) + 4
    per_point = 4 + (4 + 8 * CH) + (4 * CLASSES + 8) + 28
    return {"arm": "device", "ms_total": ms, "voted_points_per_batch": BATCH * nv,
            "algorithmic_bytes_per_batch": per_point * BATCH * nv, "achieved_gbs": per_point * BATCH * nv / ms / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--host-only", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(host_arm(a.iters)))
    if not a.host_only:
        print(json.dumps(device_arm(a.iters)))


if __name__ == "__main__":
    main()
