"""ncu launch list (gpu__time_duration.sum CSV) -> markdown table of per-kernel time shares per step.
usage: python tools/launches_summary.py gpurun_out/launches.csv STEPS_CAPTURED > profiles/<name>.md"""
import csv, re, sys
from collections import defaultdict

path, steps = sys.argv[1], float(sys.argv[2])
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    rows.append((r["Kernel Name"], us))
agg = defaultdict(lambda: [0, 0.0])
for k, us in rows:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
print("| kernel | launches/step | avg us | us/step | share |\n|---|---:|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %.1f | %.1f | %.1f | %.1f%% |" % (k[:70], n / steps, us / n, us / steps, 100 * us / tot))
print("\nTotal kernel time per step: %.2f ms (%d launches captured over %g steps)." % (tot / steps / 1e3, len(rows), steps))
