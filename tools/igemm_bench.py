"""Times the tcgen05 implicit-GEMM conv at the metric shape (B=16,R=32,C=64).  GPU box only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pvcnn_b200 import dense


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
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
hi, lo = dense.split_tf32(x, want_hi=False)
flops = 2.0 * b * r ** 3 * c * c * 27
for npass in (1, 3):
    ms = timeit(lambda: dense.igemm_conv(hi, lo, w_hi, w_lo, None, npass=npass))
    print(json.dumps({"op": "conv3d_fwd", "npass": npass, "ms": ms, "tflops_algorithmic": flops / ms / 1e9}))
g = torch.randn(b, r, r, r, c, device="cuda")
g_hi, g_lo = dense.split_tf32(g, want_hi=False)
for npass in (1, 3):
    ms = timeit(lambda: dense.conv_wgrad(hi, lo, g_hi, g_lo, c, c, 27, npass=npass))
    print(json.dumps({"op": "conv3d_wgrad", "npass": npass, "ms": ms, "tflops_algorithmic": flops / ms / 1e9}))
xp = torch.randn(1, 1, 1, b * 4096, c, device="cuda")
wp = torch.randn(c, c, 1, device="cuda")
wp_hi, wp_lo = dense.prep_weight(wp)
php, plo = dense.split_tf32(xp, want_hi=False)
ms = timeit(lambda: dense.igemm_conv(php, plo, wp_hi, wp_lo, None, npass=3))
print(json.dumps({"op": "pointwise_gemm", "npass": 3, "ms": ms}))
ms = timeit(lambda: dense.split_tf32(x, want_hi=False))
print(json.dumps({"op": "split_tf32_134MB", "ms": ms}))
torch.backends.cudnn.benchmark = True
xn = x.permute(0, 4, 1, 2, 3).contiguous()
for tf in (True, False):
    torch.backends.cudnn.allow_tf32 = tf
    ms = timeit(lambda: torch.nn.functional.conv3d(xn, w, None, padding=1))
    print(json.dumps({"op": "cudnn_conv3d_fwd", "allow_tf32": tf, "ms": ms, "tflops": flops / ms / 1e9}))
