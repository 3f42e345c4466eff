"""Stand-alone ops: ours (C ABI) vs the reference's own CUDA kernels (oracle/_ref .so) on the same B200.
Writes gpurun_out/ops_bench.json; summarised in profiles/.  GPU box only."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from pvcnn_b200.functional import backend as B
from oracle.ref_gpu import backend as ref_backend

R = ref_backend()


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    for i in range(iters):
        ev[i].record(); fn()
    ev[iters].record(); torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2] * 1e3  # us


rows = []


def add(name, shape, ours, ref, algo_bytes=None):
    t_o, t_r = timeit(ours), timeit(ref)
    row = {"op": name, "shape": shape, "ours_us": round(t_o, 1), "reference_us": round(t_r, 1), "speedup": round(t_r / t_o, 1)}
    if algo_bytes:
        row["ours_GBps_algorithmic"] = round(algo_bytes / t_o / 1e3, 1)
    rows.append(row)
    print(json.dumps(row), flush=True)


g = torch.Generator(device="cuda").manual_seed(1588147245)
# ---- metric shapes: B=16 N=4096 C=64 R=32
b, c, n, r = 16, 64, 4096, 32
f = torch.randn(b, c, n, device="cuda", generator=g)
co = torch.rand(b, 3, n, device="cuda", generator=g) * torch.tensor([1.5, 1.5, 3.0], device="cuda").view(1, 3, 1)
from pvcnn_b200 import functional as F
nc, vc = F.voxelize_coords(co, r)
add("avg_voxelize fwd", "B16 C64 N4096 R32", lambda: B.avg_voxelize_forward(f, vc, r), lambda: R.avg_voxelize_forward(f, vc, r),
    4 * b * (c * n + 3 * n + c * r ** 3 + n + r ** 3))
out, ind, cnt = B.avg_voxelize_forward(f, vc, r)
gy = torch.randn(b, c, r ** 3, device="cuda", generator=g)
add("avg_voxelize bwd", "B16 C64 N4096 R32", lambda: B.avg_voxelize_backward(gy, ind, cnt), lambda: R.avg_voxelize_backward(gy, ind, cnt),
    4 * b * (2 * c * n + 2 * n))
add("trilinear_devoxelize fwd (train)", "B16 C64 N4096 R32", lambda: B.trilinear_devoxelize_forward(r, True, nc, gy),
    lambda: R.trilinear_devoxelize_forward(r, True, nc, gy), 4 * b * (3 * n + c * min(r ** 3, 8 * n) + c * n) + 8 * b * 8 * n)
o, inds, wg = B.trilinear_devoxelize_forward(r, True, nc, gy)
go = torch.randn(b, c, n, device="cuda", generator=g)
add("trilinear_devoxelize bwd", "B16 C64 N4096 R32", lambda: B.trilinear_devoxelize_backward(go, inds, wg, r),
    lambda: R.trilinear_devoxelize_backward(go, inds, wg, r), 4 * b * (c * n + 16 * n + c * r ** 3))
# ---- PVCNN++ SA0 shapes: B=8 N=8192 M=1024 U=32
b, n, m, u, c = 8, 8192, 1024, 32, 32
p = torch.rand(b, 3, n, device="cuda", generator=g)
add("furthest_point_sampling", "B8 N8192 M1024", lambda: B.furthest_point_sampling(p, m), lambda: R.furthest_point_sampling(p, m))
fidx = B.furthest_point_sampling(p, m)
ce = B.gather_features_forward(p, fidx)
add("gather fwd", "B8 C3 N8192 M1024", lambda: B.gather_features_forward(p, fidx), lambda: R.gather_features_forward(p, fidx))
add("ball_query", "B8 N8192 M1024 r0.1 U32", lambda: B.ball_query(ce, p, 0.1, u), lambda: R.ball_query(ce, p, 0.1, u))
bq = B.ball_query(ce, p, 0.1, u)
pf = torch.randn(b, c, n, device="cuda", generator=g)
add("grouping fwd", "B8 C32 N8192 M1024 U32", lambda: B.grouping_forward(pf, bq), lambda: R.grouping_forward(pf, bq),
    4 * b * (c * m * u + m * u + c * m * u))
ggy = torch.randn(b, c, m, u, device="cuda", generator=g)
add("grouping bwd", "B8 C32 N8192 M1024 U32", lambda: B.grouping_backward(ggy, bq, n), lambda: R.grouping_backward(ggy, bq, n))
# fused BallQuery grouping (SA0 of PVCNN++: 9 -> 32 channels carried, 3 coordinate channels prepended)
def _ref_group_concat():
    rel = R.grouping_forward(p, bq) - ce.unsqueeze(-1)
    return torch.cat([rel, R.grouping_forward(pf, bq)], dim=1)
add("BallQuery grouping+centre+cat fwd (fused)", "B8 C32+3 N8192 M1024 U32", lambda: B.group_concat_forward(p, ce, pf, bq),
    _ref_group_concat, 4 * b * ((c + 3) * m * u + m * u + (c + 3) * m * u))
cgy = torch.randn(b, c + 3, m, u, device="cuda", generator=g)
def _ref_group_concat_bwd():
    return R.grouping_backward(cgy[:, 3:].contiguous(), bq, n)
add("BallQuery grouping+centre+cat bwd (features only)", "B8 C32+3 N8192 M1024 U32",
    lambda: B.group_concat_backward(cgy, bq, n), _ref_group_concat_bwd)
cf = torch.randn(b, 64, m, device="cuda", generator=g)
add("three_nn_interpolate fwd", "B8 C64 N8192 M1024", lambda: B.three_nearest_neighbors_interpolate_forward(p, ce, cf),
    lambda: R.three_nearest_neighbors_interpolate_forward(p, ce, cf))
io, ii, iw = B.three_nearest_neighbors_interpolate_forward(p, ce, cf)
igy = torch.randn(b, 64, n, device="cuda", generator=g)
add("three_nn_interpolate bwd", "B8 C64 N8192 M1024", lambda: B.three_nearest_neighbors_interpolate_backward(igy, ii, iw, m),
    lambda: R.three_nearest_neighbors_interpolate_backward(igy, ii, iw, m))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ops_bench.json"), "w"), indent=1)
