#!/bin/bash
mkdir -p gpurun_out
export PVCNN_BENCH_MINIMAL=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/launches_run.log 2>&1
tail -2 gpurun_out/launches_run.log | cut -c1-200
