"""Debug aid: fused and composed PVConv at the full metric configuration against the fp64 CPU oracle."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import oracle
from util import rng, s3dis_like_coords, rel_err
from test_pvconv_gpu import make_block, _step

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
b, n, c, r = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (16, 4096, 64, 32)))
g = rng(1588147245 % (2 ** 31))
f = g.standard_normal((b, c, n), dtype=np.float32)
co = s3dis_like_coords(g, b, n)
go = g.standard_normal((b, c, n), dtype=np.float32)
m = make_block(c, c, r).cuda().train()
params = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items() if "running" not in k and "num_batches" not in k}
t0 = time.time()
ref = oracle.pvconv_forward_backward(params, f, co, go, r, training=True, dtype="float64", buffers=None, threads=os.cpu_count())
print("oracle fp64: %.1f s; keys %s" % (time.time() - t0, sorted(ref.keys())[:6]), flush=True)
ft, cot, got = (torch.from_numpy(a).cuda() for a in (f, co, go))
def l2(a, b_):
    a, b_ = np.asarray(a, np.float64), np.asarray(b_, np.float64)
    return float(np.linalg.norm(a - b_) / max(np.linalg.norm(b_), 1e-30))
for mode in ("composed", "fused"):
    os.environ["PVCNN_B200_PVCONV"] = mode
    out, gin, pg = _step(m, ft, cot, got)
    print("%-9s out max %.2e l2 %.2e | gin max %.2e l2 %.2e" % (mode, rel_err(out.cpu().numpy(), ref["out"]), l2(out.cpu().numpy(), ref["out"]),
          rel_err(gin.cpu().numpy(), ref["grad_features"]), l2(gin.cpu().numpy(), ref["grad_features"])), flush=True)
    for k, v in pg.items():
        kk = "grad_" + k if ("grad_" + k) in ref else k
        src = ref["grads"][k] if "grads" in ref and k in ref["grads"] else ref.get(kk)
        if src is None:
            continue
        print("   %-40s l2 %.2e  (|ref| max %.2e)" % (k, l2(v.cpu().numpy(), src), np.abs(src).max()))
    res_gin = gin.cpu().numpy()
    d = np.abs(res_gin.astype(np.float64) - ref["grad_features"])
    top = np.argsort(d.ravel())[-5:][::-1]
    print("   top gin errors at (b,c,i):", [tuple(int(v) for v in np.unravel_index(t, d.shape)) for t in top], "values", d.ravel()[top], flush=True)
    per_b = np.sqrt((d ** 2).sum(axis=(1, 2)))
    print("   per-sample err norm:", np.array2string(per_b, precision=3))
    np.save(os.path.join("gpurun_out", "gin_%s.npy" % mode), res_gin)
np.save(os.path.join("gpurun_out", "gin_oracle.npy"), ref["grad_features"].astype(np.float32))
import modules
from pvcnn_b200.functional import ops as _ops
vc = _ops.voxelize_coords(cot, r, True, 0.0)[1]
if vc is not None:
    print("vox coords equal to oracle:", bool((vc.cpu().numpy() == ref["vox_coords"]).all()), "mismatches", int((vc.cpu().numpy() != ref["vox_coords"]).sum()))
np.save(os.path.join("gpurun_out", "vox_oracle.npy"), ref["vox_coords"].astype(np.int8))
