"""A/B of the weight-gradient kernels at the metric shape: dz-merged (conv_wgrad_dz.cu) vs per-tap (conv_wgrad.cu)."""
import json, os, sys
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


for (b, r, cin, cout) in [(16, 32, 64, 64), (16, 16, 64, 128), (8, 8, 256, 256)]:
    x = torch.randn(b, r, r, r, cin, device="cuda")
    g = torch.randn(b, r, r, r, cout, device="cuda")
    x_lo = dense.split_tf32(x, want_hi=False)[1]
    g_lo = dense.split_tf32(g, want_hi=False)[1]
    ref = None
    for ver in ("v1", "dz"):
        os.environ["PVCNN_B200_WGRAD"] = ver
        for npass in (3, 1):
            dw = dense.conv_wgrad(x, x_lo, g, g_lo, cin, cout, 27, npass=npass)
            if ref is None:
                ref = dw.clone()
            ms = timeit(lambda: dense.conv_wgrad(x, x_lo, g, g_lo, cin, cout, 27, npass=npass))
            print(json.dumps({"shape": [b, r, cin, cout], "kernel": ver, "npass": npass, "ms": round(ms, 4),
                              "tflops": round(2.0 * b * r ** 3 * cin * cout * 27 / ms / 1e9, 1),
                              "diff_vs_v1_3pass": float((dw - ref).abs().max() / ref.abs().max())}), flush=True)
