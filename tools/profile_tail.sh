#!/bin/bash
# ncu --set full of the HBM-bound kernels of the fused step (scatter / gather / streaming passes) and of the wgrad
# launches, taken from the 4th step of the bench loop: achieved DRAM bytes and duration per launch (north-star: "each
# kernel ships with an ncu capture reporting achieved HBM GB/s").  Numbers printed by this run are NOT bench values.
mkdir -p gpurun_out
export PVCNN_BENCH_MINIMAL=1
ncu --set full --clock-control none --import-source on \
    -k regex:'voxelize_cl_kernel|devox_fused_kernel|bwd_points_kernel|bn_apply_leaky_kernel|class_colsum_kernel|bn_bwd_apply_kernel|conv_wgrad' \
    -s 30 -c 10 -f -o gpurun_out/r02_tail python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_tail.log 2>&1
tail -2 gpurun_out/ncu_tail.log | cut -c1-200
