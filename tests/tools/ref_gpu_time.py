"""Times the reference PVConv (reference CUDA ops + cuDNN) on the metric config.  GPU box only."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle.ref_gpu import RefPVConv


def run(allow_tf32, steps=30, warmup=10, B=16, N=4096, C=64, R=32):
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    g = torch.Generator(device="cuda").manual_seed(1588147245)
    m = RefPVConv(C, C, 3, R).cuda().train()
    f = torch.randn(B, C, N, device="cuda", generator=g).requires_grad_(True)
    co = torch.rand(B, 3, N, device="cuda", generator=g) * torch.tensor([1.5, 1.5, 3.0], device="cuda").view(1, 3, 1)
    go = torch.randn(B, C, N, device="cuda", generator=g)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    for i in range(warmup + steps):
        if i >= warmup:
            ev[i - warmup].record()
        for p in m.parameters():
            p.grad = None
        f.grad = None
        out, _ = m((f, co))
        out.backward(go)
    ev[steps].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    med = ts[len(ts) // 2]
    return {"allow_tf32": allow_tf32, "ms_per_step_median": med, "points_per_s": B * N / med * 1e3}


if __name__ == "__main__":
    for tf in (True, False):
        print(json.dumps(run(tf)))
