import os, sys, json, ctypes
os.environ["PVCNN_STALL_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, modules
from pvcnn_b200 import _lib
lib = _lib.load()
torch.manual_seed(bench.SEED)
m = modules.PVConv(64, 64, 3, 32).cuda().train()
f, c, g = [t.cuda() for t in bench.make_inputs(torch, None)]
for sparse in ("1", "0"):
    os.environ["PVCNN_B200_SPARSE"] = sparse
    with torch.no_grad():
        for _ in range(3):
            m((f, c))     # last halo kernel of a forward = conv2 forward
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 8)()
    lib.pvcnn_stall_profile_read(buf)
    v = list(buf)
    print(json.dumps({"sparse": sparse, "kernel": "conv2 fwd (halo)", "units_cta0": v[4], "mma_total": v[3],
                      "mma_stall_a_ready": v[0], "mma_stall_b_full": v[1], "mma_stall_acc_empty": v[2],
                      "epi_stall_acc_full": v[5], "epi_total": v[6],
                      "cycles_per_unit": v[3] / max(1, v[4])}))
