import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, modules
from pvcnn_b200 import fused
torch.manual_seed(bench.SEED)
m = modules.PVConv(64, 64, 3, 32).cuda().train()
f, c, g = [t.cuda() for t in bench.make_inputs(torch, None)]
f.requires_grad_(True)
orig = fused._PVConvFused.forward
plans = []
def fwd(ctx, *a):
    out = orig(ctx, *a); plans.append(ctx.plan); return out
fused._PVConvFused.forward = staticmethod(fwd)
out, _ = m((f, c))
torch.cuda.synchronize()
cnt = plans[0].t["sparse"][:8].cpu().tolist()
print("counts [fwd1 units, dgrad1 units, fwd2 units, wg1 ktiles, wg2 ktiles, dg2 units]:", cnt, "of units", 16*16*8, "ktiles", 16*32*32)
