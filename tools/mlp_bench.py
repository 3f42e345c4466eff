"""Times one frozen SharedMLP layer (1x1 conv + BatchNorm + ReLU) at the layer shapes of the reference networks:
the layer-by-layer path (GEMM -> y, BatchNorm/ReLU pass -> z, z_lo), the fused-epilogue path, and torch's cuBLAS GEMM alone
(TF32 and fp32) as the library yardstick.  GPU box only.   python tools/mlp_bench.py [fp32|tf32]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pvcnn_b200 import _lib, mlp  # noqa: E402

LL = ctypes.c_longlong
npass = 1 if (len(sys.argv) > 1 and sys.argv[1] == "tf32") else 3


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


lib = mlp._lib_sizes()
shapes = [(65536, 128, 1024), (65536, 1600, 512), (65536, 512, 256), (32768, 2307, 512), (32768, 128, 1024),
          (65536, 64, 64), (262144, 67, 64), (262144, 64, 128), (131072, 259, 256)]
if os.environ.get("MLP_BENCH_SHAPE"):
    shapes = [shapes[int(os.environ["MLP_BENCH_SHAPE"])]]
print("npass=%d   rows cin cout | layered us | fused us | cuBLAS tf32 us | cuBLAS fp32 us | fused TFLOP/s (algorithmic)" % npass)
for rows, cin, cout in shapes:
    dev = "cuda"
    ci, co = mlp._pad4(cin), mlp._pad4(cout)
    x = torch.randn(rows, ci, device=dev)
    xl = torch.randn(rows, ci, device=dev) * 1e-4
    w = torch.randn(cout, cin, 1, device=dev) * 0.05
    bias, gamma, beta = torch.randn(cout, device=dev), torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
    rm, rv = torch.randn(cout, device=dev) * 0.1, torch.rand(cout, device=dev) + 0.5
    wprep = torch.empty(lib.pvcnn_mlp_wprep_floats(cin, cout), device=dev)
    partials = torch.empty(lib.pvcnn_mlp_partials_floats(cout), device=dev)
    coef = torch.empty(4 * co, device=dev)
    y, z, zl = (torch.empty(rows, co, device=dev) for _ in range(3))
    _lib.call("pvcnn_mlp_layer_prepare", cin, cout, 1e-5, w, gamma, beta, rm, rv, wprep, coef)
    t_lay = timeit(lambda: _lib.call("pvcnn_mlp_layer_forward", LL(rows), cin, cout, 0, npass, 1e-5, 0.1, x, xl, w, bias, gamma,
                                     beta, rm, rv, None, wprep, partials, coef, y, z, zl, 0, None, None, None, LL(0), None, 0))
    z1 = z.clone()
    t_fus = timeit(lambda: _lib.call("pvcnn_mlp_layer_forward_eval", LL(rows), cin, cout, npass, x, xl, wprep, bias, coef, LL(0),
                                     None, 0, None, z, zl, 0, None, None, None))
    same = torch.equal(z, z1)
    w2 = w[:, :, 0].t().contiguous()
    xx = x[:, :cin].contiguous()
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf = timeit(lambda: torch.mm(xx, w2))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_32 = timeit(lambda: torch.mm(xx, w2))
    print("%7d %5d %5d | %8.1f | %8.1f | %8.1f | %8.1f | %6.1f  %s" % (rows, cin, cout, t_lay, t_fus, t_tf, t_32,
                                                                     2.0 * rows * cin * cout / t_fus / 1e6, "" if same else "MISMATCH"), flush=True)
