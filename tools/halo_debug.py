import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pvcnn_b200 import dense

def run(b, r, cin, cout, npass):
    torch.manual_seed(2)
    x = torch.randn(b, cin, r, r, r, device="cuda")
    conv = torch.nn.Conv3d(cin, cout, 3, padding=1).cuda()
    ref = torch.nn.functional.conv3d(x.double().cpu(), conv.weight.double().cpu(), conv.bias.double().cpu(), padding=1)
    xcl = x.permute(0, 2, 3, 4, 1).contiguous()
    a_hi, a_lo = dense.split_tf32(xcl, want_hi=False)
    w_hi, w_lo = dense.prep_weight(conv.weight)
    outs = {}
    for mode in ("v1", "halo"):
        os.environ["PVCNN_B200_CONV"] = mode
        o = dense.igemm_conv(a_hi, a_lo, w_hi, w_lo, conv.bias.detach(), npass=npass)
        outs[mode] = o[..., :cout].permute(0, 4, 1, 2, 3).double().cpu()
    scale = ref.abs().max()
    for mode, o in outs.items():
        err = (o - ref).abs() / scale
        print(f"cfg b{b} r{r} ci{cin} co{cout} npass{npass} {mode}: max {err.max():.2e} mean {err.mean():.2e}")
        if mode == "halo" and err.max() > 2e-5:
            bad = err > 2e-5
            print("   bad fraction", bad.float().mean().item())
            e = err  # [b, co, x, y, z]
            print("   by x:", [f"{v:.1e}" for v in e.amax(dim=(0, 1, 3, 4)).tolist()])
            print("   by y:", [f"{v:.1e}" for v in e.amax(dim=(0, 1, 2, 4)).tolist()])
            print("   by z:", [f"{v:.1e}" for v in e.amax(dim=(0, 1, 2, 3)).tolist()])
            print("   by co:", [f"{v:.1e}" for v in e.amax(dim=(0, 2, 3, 4)).tolist()][:16])
            print("   by b:", [f"{v:.1e}" for v in e.amax(dim=(1, 2, 3, 4)).tolist()])

for cfg in [(2, 8, 16, 16), (1, 16, 64, 64), (2, 32, 64, 64), (1, 16, 9, 64), (1, 8, 32, 32)]:
    for npass in (1, 3):
        run(*cfg, npass)
