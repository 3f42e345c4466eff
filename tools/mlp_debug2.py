import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import modules
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
b, cin, widths, n = 4, 1472, [512, 256], 512
torch.manual_seed(seed)
g = np.random.default_rng(seed)
prod = modules.SharedMLP(cin, widths, dim=1).cuda().train()
x = g.standard_normal((b, cin, n), dtype=np.float32)
go = g.standard_normal((b, widths[-1], n), dtype=np.float32)
res = {}
for arm in ("native", "torch", "native"):
    os.environ["PVCNN_B200_MLP"] = arm
    for p in prod.parameters(): p.grad = None
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    out = prod(xt); out.backward(torch.from_numpy(go).cuda())
    torch.cuda.synchronize()
    res.setdefault(arm, []).append({k: p.grad.clone() for k, p in prod.named_parameters()})
for k in res["torch"][0]:
    a, b2, t = res["native"][0][k], res["native"][1][k], res["torch"][0][k]
    print(k, "native1-vs-torch %.2e  native2-vs-torch %.2e  native1-vs-native2 %.2e" % (
        float((a - t).abs().max() / t.abs().max().clamp_min(1e-30)), float((b2 - t).abs().max() / t.abs().max().clamp_min(1e-30)),
        float((a - b2).abs().max() / t.abs().max().clamp_min(1e-30))))
bad = (res["native"][0]["layers.4.bias"] - res["torch"][0]["layers.4.bias"]).abs()
print("layer-2 dbeta: worst channels", torch.topk(bad, 5).indices.tolist(), torch.topk(bad, 5).values.tolist())
