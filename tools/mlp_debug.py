import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import modules
from oracle import ref_modules as R
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for (b, cin, widths, n) in [(4, 1472, [512, 256], 512), (2, 9, [64, 64], 1024)]:
    for seed in range(6):
        torch.manual_seed(seed)
        g = np.random.default_rng(seed)
        prod = modules.SharedMLP(cin, widths, dim=1)
        ref = R.clone_as_oracle(prod, R.SharedMLP(cin, widths, dim=1))
        prod = prod.cuda().train()
        x = g.standard_normal((b, cin, n), dtype=np.float32)
        go = g.standard_normal((b, widths[-1], n), dtype=np.float32)
        xr = torch.from_numpy(x).double().requires_grad_(True)
        outr = ref(xr); outr.backward(torch.from_numpy(go).double())
        res = {}
        for arm in ("native", "torch"):
            os.environ["PVCNN_B200_MLP"] = arm
            for p in prod.parameters(): p.grad = None
            xt = torch.from_numpy(x).cuda().requires_grad_(True)
            out = prod(xt); out.backward(torch.from_numpy(go).cuda())
            pe = {k: rel(p.grad.cpu().numpy().reshape(dict(ref.named_parameters())[k].shape), dict(ref.named_parameters())[k].grad.numpy()) for k, p in prod.named_parameters()}
            res[arm] = (rel(out.detach().cpu().numpy(), outr.detach().numpy()), rel(xt.grad.cpu().numpy(), xr.grad.numpy()), pe)
        print(cin, widths, "seed", seed)
        for arm, (eo, eg, pe) in res.items():
            print("   %-6s out %.2e  xgrad %.2e  " % (arm, eo, eg) + " ".join("%s=%.1e" % (k.replace("layers.", ""), v) for k, v in pe.items()))
