"""Debug aid: is the run-to-run spread of block 2's point-branch weight gradient inside the network explained by its own
inputs (conditioning) or by the kernels?  Each run is compared with an fp64 evaluation on ITS OWN captured inputs."""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
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
blk = net.blocks[2]
conv, bn = blk.point_features.layers[0], blk.point_features.layers[1]

def run(mode):
    os.environ["PVCNN_B200_PVCONV"] = mode
    net.load_state_dict(state); net.zero_grad(set_to_none=True)
    cap = {}
    def _hook(m, i, o):
        o[0].retain_grad()
        cap.update(fin=i[0][0].detach().clone(), out=o[0])
    h = blk.register_forward_hook(_hook)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    F.cross_entropy(net(xt), labels).backward()
    h.remove()
    gout = cap["out"].grad.detach().clone()
    wp = conv.weight.grad.detach().clone()
    w = conv.weight.detach().double().requires_grad_(True)
    y = F.conv1d(cap["fin"].double(), w, conv.bias.detach().double())
    z = F.batch_norm(y, None, None, bn.weight.detach().double(), bn.bias.detach().double(), True, 0.1, bn.eps)
    F.relu(z).backward(gout.double())
    return dict(fin=cap["fin"].cpu().numpy(), gout=gout.cpu().numpy(), wp=wp.cpu().numpy(), wp64=w.grad.cpu().numpy())

R = {"c": run("composed"), "f1": run("fused"), "f2": run("fused")}
for k, r in R.items():
    print("%-3s wp vs own fp64 %.2e | fin vs c %.2e | gout vs c %.2e | wp64 vs c's wp64 %.2e | wp vs c %.2e" % (
        k, rel_err(r["wp"], r["wp64"]), rel_err(r["fin"], R["c"]["fin"]), rel_err(r["gout"], R["c"]["gout"]),
        rel_err(r["wp64"], R["c"]["wp64"]), rel_err(r["wp"], R["c"]["wp"])))
d = R["f1"]["gout"] - R["c"]["gout"]
print("gout diff f1-c: per-channel mean of diff / abs mean of gout:", float(np.abs(d.mean(axis=(0, 2))).max() / np.abs(R["c"]["gout"]).mean()))
