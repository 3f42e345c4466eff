"""Launches the dense halo conv kernel a few times in each precision for an `ncu --set full` capture
(tools/profile_halo.sh).  Not a benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pvcnn_b200 import dense
b, r, c = 16, 32, 64
x = torch.randn(b, r, r, r, c, device="cuda")
w = torch.randn(c, c, 3, 3, 3, device="cuda") * 0.05
w_hi, w_lo = dense.prep_weight(w)
x_lo = dense.split_tf32(x, want_hi=False)[1]
for npass in (3, 1):
    for _ in range(3):
        dense.igemm_conv(x, x_lo, w_hi, w_lo, None, npass=npass)
torch.cuda.synchronize()
