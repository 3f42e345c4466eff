#!/bin/bash
# ncu evidence for the bench command: (1) launch list with per-launch device time, (2) full capture of the
# dominant kernels.  Numbers printed by these runs are NOT bench values.
mkdir -p gpurun_out
export PVCNN_BENCH_MINIMAL=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv_halo_kernel|conv_wgrad_kernel' -s 12 -c 4 \
    -o gpurun_out/prof_dense python bench.py --steps 1 --warmup 3 > gpurun_out/prof_dense.log 2>&1
ls -la gpurun_out/
