"""Debug aid: chained fused PVConv blocks -- verify the activity lists of every block stay intact until its backward."""
import os, sys
import numpy as np
import torch
import torch.nn as nn
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import modules
from pvcnn_b200 import fused, _lib
from util import rng, s3dis_like_coords
from test_network_gpu import MiniPVCNN

os.environ["PVCNN_B200_PVCONV"] = "fused"
plans = []
orig_init = fused._Plan.__init__
def init(self, desc, device, need_backward):
    orig_init(self, desc, device, need_backward)
    plans.append(self)
fused._Plan.__init__ = init

snap = {}
orig_call = _lib.call
def call(name, *a, **k):
    if name == "pvcnn_pvconv_backward":
        torch.cuda.synchronize()
        for i, p in enumerate(plans):
            cur = p.t["sparse"].cpu().numpy()
            d = np.nonzero(cur != snap[i])[0]
            print(f"before backward: plan {i} (r={p.desc.r} cin={p.desc.cin} cout={p.desc.cout}) sparse diffs: {len(d)}"
                  + (f" first at int {d[0]}..{d[-1]} of {cur.size}" if len(d) else ""), flush=True)
    r = orig_call(name, *a, **k)
    if name == "pvcnn_pvconv_backward":
        torch.cuda.synchronize()
        print("  backward ok", flush=True)
    return r
_lib.call = call
fused._lib.call = call

g = rng(50)
b, n = 4, 2048
x = np.concatenate([s3dis_like_coords(g, b, n), g.random((b, 6, n), dtype=np.float32)], axis=1)
labels = torch.from_numpy(g.integers(0, 13, size=(b, n))).cuda()
torch.manual_seed(3)
net = MiniPVCNN().cuda().train()
# dirty the allocator so torch.empty returns garbage
junk = [torch.full((1 << 22,), 1e-3 * (i + 1), device="cuda") for i in range(8)]
del junk
xt = torch.from_numpy(x).cuda().requires_grad_(True)
logits = net(xt)
torch.cuda.synchronize()
for i, p in enumerate(plans):
    snap[i] = p.t["sparse"].cpu().numpy().copy()
    c = snap[i][:8]
    print(f"plan {i}: r={p.desc.r} counts={c.tolist()} total ints {snap[i].size}", flush=True)
loss = nn.functional.cross_entropy(logits, labels)
loss.backward()
torch.cuda.synchronize()
print("done")
