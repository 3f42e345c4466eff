"""A/B of the halo conv kernels at the metric shape (dense walk): v2 vs v3, 3xTF32 and single-pass TF32, plus the
in-kernel stall counters of CTA 0 (PVCNN_STALL_PROFILE=1).  python tools/halo_bench.py"""
import ctypes, json, os, sys
os.environ["PVCNN_STALL_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pvcnn_b200 import dense, _lib


def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    for i in range(iters):
        ev[i].record(); fn()
    ev[iters].record(); torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


lib = _lib.load()
shapes = [(16, 32, 64, 64)] + ([(16, 16, 64, 128), (8, 8, 256, 256), (32, 12, 64, 128)] if "--all" in sys.argv else [])
for (b, r, cin, cout) in shapes:
    x = torch.randn(b, r, r, r, cin, device="cuda")
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
    w_hi, w_lo = dense.prep_weight(w)
    x_lo = dense.split_tf32(x, want_hi=False)[1]
    ref = None
    for ver in ("v3", "v1"):
        os.environ["PVCNN_B200_CONV"] = ver.split("/")[0]
        for npass in (3, 1):
            try:
                out = dense.igemm_conv(x, x_lo, w_hi, w_lo, None, npass=npass)
            except Exception as e:  # outside the kernel's envelope
                print(json.dumps({"shape": [b, r, cin, cout], "kernel": ver, "npass": npass, "error": str(e)[:80]}))
                continue
            if ref is None and npass == 3:
                ref = out.clone()
            err = float((out - ref).abs().max() / ref.abs().max()) if ref is not None else None
            ms = timeit(lambda: dense.igemm_conv(x, x_lo, w_hi, w_lo, None, npass=npass))
            buf = (ctypes.c_longlong * 8)()
            lib.pvcnn_stall_profile_read(buf)
            v = list(buf)
            gflop = 2.0 * b * r ** 3 * cin * cout * 27 / 1e9
            print(json.dumps({"shape": [b, r, cin, cout], "kernel": ver, "npass": npass, "ms": round(ms, 4),
                              "tflops": round(gflop / ms, 1), "diff_vs_first": err,
                              "cta0": {"items": v[4], "mma_total": v[3], "stall_a": v[0], "stall_b": v[1],
                                       "stall_acc": v[2], "epi_stall_full": v[5], "epi_total": v[6]}}), flush=True)
