import os, sys, json, ctypes
os.environ["PVCNN_STALL_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pvcnn_b200 import dense, _lib
b, r, c = 16, 32, 64
x = torch.randn(b, r, r, r, c, device="cuda")
g = torch.randn(b, r, r, r, c, device="cuda")
lib = _lib.load()
x_lo = dense.split_tf32(x, want_hi=False)[1]
g_lo = dense.split_tf32(g, want_hi=False)[1]
import time
for G, lo in (("1", "k"), ("1", "g"), ("2", "k"), ("2", "g")):
    npass = 3
    os.environ["PVCNN_WGRAD_G"] = G
    os.environ["PVCNN_WGRAD_LO"] = lo
    for _ in range(3):
        dense.conv_wgrad(x, x_lo, g, g_lo, c, c, 27, npass=npass)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dense.conv_wgrad(x, x_lo, g, g_lo, c, c, 27, npass=npass)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"G": G, "lo": lo, "ms": e0.elapsed_time(e1) / 5}))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 8)()
    lib.pvcnn_stall_profile_read(buf)
    v = list(buf)
    print(json.dumps({"kernel": "wgrad", "npass": npass, "producer_stall_on_empty": v[0], "producer_total": v[1],
                      "mma_stall_on_ready": v[2], "mma_stall_on_tmem_empty": v[3], "mma_total": v[4],
                      "converter_stall_on_full": v[5], "converter_total": v[6], "converter_fence": v[7]}))
