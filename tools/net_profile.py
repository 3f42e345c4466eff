"""Per-kernel GPU time of one eval forward of a BASELINE network on the native path (torch profiler / CUPTI).
python tools/net_profile.py frustum_pvcnne|s3dis_pvcnn|pvcnn2 [tf32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from pvcnn_b200 import zoo
cfg = sys.argv[1]
if len(sys.argv) > 2: os.environ["PVCNN_B200_PRECISION"] = sys.argv[2]
torch.manual_seed(0)
model, spec = zoo.build(cfg)
model = model.cuda().eval()
x = zoo.synthetic_input(spec, torch.Generator().manual_seed(1))
x = {k: v.cuda() for k, v in x.items()} if isinstance(x, dict) else x.cuda()
np.random.seed(0)
with torch.no_grad():
    for _ in range(3): model(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): model(x)
        torch.cuda.synchronize()
rows = [(e.key, e.count / 3, e.device_time_total / 3) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print("%s: %.3f ms of kernel time per forward" % (cfg, tot / 1e3))
for k, n, t in rows[:22]:
    print("  %-70s %5.1f x  %8.1f us  %4.1f%%" % (k[:70], n, t, 100 * t / tot))
