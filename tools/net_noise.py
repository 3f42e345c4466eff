"""Debug aid: run-to-run and mode-to-mode differences of the mini network gradients."""
import os, sys
import numpy as np
import torch
import torch.nn as nn
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from util import rng, s3dis_like_coords, rel_err
from test_network_gpu import MiniPVCNN

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = rng(50)
b, n = 4, 2048
x = np.concatenate([s3dis_like_coords(g, b, n), g.random((b, 6, n), dtype=np.float32)], axis=1)
labels = torch.from_numpy(g.integers(0, 13, size=(b, n))).cuda()
torch.manual_seed(3)
net = MiniPVCNN().cuda().train()
state = {k: v.clone() for k, v in net.state_dict().items()}
def run(mode, sparse="1"):
    os.environ["PVCNN_B200_PVCONV"] = mode
    os.environ["PVCNN_B200_SPARSE"] = sparse
    net.load_state_dict(state); net.zero_grad(set_to_none=True)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    feats = []
    hooks = [blk.register_forward_hook(lambda m, i, o: (o[0] if isinstance(o, tuple) else o).retain_grad() or feats.append(o[0] if isinstance(o, tuple) else o)) for blk in net.blocks]
    loss = nn.functional.cross_entropy(net(xt), labels); loss.backward()
    for h in hooks: h.remove()
    r = {"x": xt.grad.cpu().numpy()}
    for i, f in enumerate(feats): r[f"dfeat{i}"] = f.grad.cpu().numpy()
    r.update({k: p.grad.cpu().numpy() for k, p in net.named_parameters()})
    return r
runs = {"c1": run("composed"), "c2": run("composed"), "f1": run("fused"), "f2": run("fused"), "fd": run("fused", "0")}
keys = ["x", "dfeat0", "dfeat1", "dfeat2", "dfeat3"] + [k for k in runs["c1"] if k.startswith("blocks") and "weight" in k]
print("%-40s %10s %10s %10s %10s" % ("tensor", "c1-c2", "f1-f2", "f1-c1", "fd-c1"))
for k in keys:
    ref = runs["c1"][k]
    if np.abs(ref).max() < 1e-7: continue
    print("%-40s %10.2e %10.2e %10.2e %10.2e" % (k, rel_err(runs["c2"][k], ref), rel_err(runs["f2"][k], runs["f1"][k]),
                                         rel_err(runs["f1"][k], ref), rel_err(runs["fd"][k], ref)))
