"""Debug aid: isolate the third block of the mini network (64->64, R=8) with its real inputs; compare the point-branch
weight gradient of fused / composed runs with an fp64 evaluation."""
import os, sys
import numpy as np
import torch
import torch.nn as nn
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
os.environ["PVCNN_B200_PVCONV"] = "composed"
cap = {}
blk = net.blocks[2]
def _hook(m, i, o):
    o[0].retain_grad()
    cap.update(fin=i[0][0].detach().clone(), coords=i[0][1].detach().clone(), out=o[0])
h1 = blk.register_forward_hook(_hook)
xt = torch.from_numpy(x).cuda().requires_grad_(True)
loss = F.cross_entropy(net(xt), labels); loss.backward()
h1.remove()
gout = cap["out"].grad.detach().clone()
fin, coords = cap["fin"], cap["coords"].contiguous()
print("block input stats: mean %.3g std %.3g max %.3g; grad_out abs mean %.3g" % (fin.mean(), fin.std(), fin.abs().max(), gout.abs().mean()))

def run(mode, prec="fp32"):
    os.environ["PVCNN_B200_PVCONV"] = mode
    os.environ["PVCNN_B200_PRECISION"] = prec
    net.load_state_dict(state); net.zero_grad(set_to_none=True)
    f = fin.clone().requires_grad_(True)
    o, _ = blk((f, coords))
    o.backward(gout)
    return dict(wp=blk.point_features.layers[0].weight.grad.cpu().numpy().copy(), gp=blk.point_features.layers[1].weight.grad.cpu().numpy().copy(),
                dx=f.grad.cpu().numpy().copy(), w2=blk.voxel_layers[3].weight.grad.cpu().numpy().copy())

# fp64 truth for the point branch (its gradient does not depend on the voxel branch: out = voxel + point)
net.load_state_dict(state)
conv, bn = blk.point_features.layers[0], blk.point_features.layers[1]
w = conv.weight.detach().double().requires_grad_(True)
f64 = fin.double().requires_grad_(True)
y = F.conv1d(f64, w, conv.bias.detach().double())
z = F.batch_norm(y, None, None, bn.weight.detach().double(), bn.bias.detach().double(), True, 0.1, bn.eps)
F.relu(z).backward(gout.double())
wp64 = w.grad.cpu().numpy()
dxp64 = f64.grad.cpu().numpy()
# conditioning: sum |terms| / |result|
print("fp64 wp grad: max %.3g" % np.abs(wp64).max())
runs = [("c", run("composed")), ("f1", run("fused")), ("f2", run("fused")), ("t", run("fused", "tf32"))]
for name, rr in runs:
    print("%-3s wp vs fp64 %.2e | wp vs c %.2e | dx vs c %.2e | gp vs c %.2e | w2 vs c %.2e" % (
        name, rel_err(rr["wp"], wp64), rel_err(rr["wp"], runs[0][1]["wp"]), rel_err(rr["dx"], runs[0][1]["dx"]),
        rel_err(rr["gp"], runs[0][1]["gp"]), rel_err(rr["w2"], runs[0][1]["w2"])))
