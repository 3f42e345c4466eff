"""Times furthest point sampling: the 4-CTA cluster kernel against the single-CTA kernel (PVCNN_B200_FPS=cta)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvcnn_b200.functional import backend as B  # noqa: E402


def timeit(fn, it=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3


for b, n, m in [(8, 8192, 1024), (8, 4096, 1024), (32, 2048, 512), (32, 1024, 256), (8, 16384, 1024), (37, 8192, 1024)]:
    p = torch.rand(b, 3, n, device="cuda")
    os.environ["PVCNN_B200_FPS"] = "t1024"
    i0 = B.furthest_point_sampling(p, m)
    line = f"B{b} N{n} M{m}:"
    for mode in ["t1024", "cta", "c1024", "c256", "c128"]:
        os.environ["PVCNN_B200_FPS"] = mode
        i1 = B.furthest_point_sampling(p, m)
        t1 = timeit(lambda: B.furthest_point_sampling(p, m))
        line += f"  {mode} {t1:7.1f}us{'' if torch.equal(i0, i1) else ' MISMATCH'}"
    print(line, flush=True)
