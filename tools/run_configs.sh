#!/bin/bash
# BASELINE.json configs 2-5 on one GPU, both precision modes -> gpurun_out/cfgF_*.json (summarised in profiles/r02_configs.md)
mkdir -p gpurun_out
for c in frustum_pvcnne pvcnn2 s3dis_pvcnn shapenet_c0p25_train; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/cfgF_$c.json 2> gpurun_out/cfgF_$c.err
done
for c in frustum_pvcnne pvcnn2 s3dis_pvcnn; do
  python bench.py --config $c --steps 10 --warmup 3 --precision tf32 > gpurun_out/cfgF_tf32_$c.json 2> gpurun_out/cfgF_tf32_$c.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/cfgF_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    g = (d.get("cuda_graph") or {}).get("ms_per_step")
    print("%-44s %-12s ours %.3f ms  graph %s  comparison %.3f  launches %d  clocks %s" % (
        f, d["dtype"], d["ms_per_step"], ("%.3f" % g) if g else "-", d["comparison_arm"]["ms_per_step"], d["gpu_launches"],
        d["clocks"].get("sm_mhz")))
PY
