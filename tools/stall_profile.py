import os, sys, json, ctypes
os.environ["PVCNN_STALL_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pvcnn_b200 import dense, _lib
b, r, c = 16, 32, 64
x = torch.randn(b, r, r, r, c, device="cuda")
g = torch.randn(b, r, r, r, c, device="cuda")
lib = _lib.load()
for npass in (1, 3):
    for _ in range(3):
        dense.conv_wgrad(x, x, g, g, c, c, 27, npass=npass)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 8)()
    lib.pvcnn_stall_profile_read(buf)
    v = list(buf)
    print(json.dumps({"kernel": "wgrad", "npass": npass, "producer_stall_on_empty": v[0], "producer_total": v[1],
                      "mma_stall_on_ready": v[2], "mma_stall_on_tmem_empty": v[3], "mma_total": v[4],
                      "converter_stall_on_full": v[5], "converter_total": v[6]}))
