import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pvcnn_b200 import dense
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    for i in range(iters):
        ev[i].record(); fn()
    ev[iters].record(); torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]
b, r, c = 16, 32, 64
x = torch.randn(b, r, r, r, c, device="cuda")
w = torch.randn(c, c, 3, 3, 3, device="cuda") * 0.05
w_hi, w_lo = dense.prep_weight(w)
ref = None
for exp in (0, 2, 0, 2):
    os.environ["PVCNN_HALO_EXP"] = str(exp)
    out = dense.igemm_conv(x, x, w_hi, w_lo, None, npass=3)
    if exp == 0: ref = out.clone()
    err = float((out - ref).abs().max() / ref.abs().max())
    ms = timeit(lambda: dense.igemm_conv(x, x, w_hi, w_lo, None, npass=3))
    print(json.dumps({"exp": exp, "ms": ms, "diff_vs_exp0": err}))
os.environ["PVCNN_HALO_EXP"] = "0"
ms = timeit(lambda: dense.igemm_conv(x, x, w_hi, w_lo, None, npass=1))
print(json.dumps({"npass": 1, "ms": ms}))
