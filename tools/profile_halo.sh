#!/bin/bash
# ncu --set full of the dense conv_halo_kernel launches (3xTF32 then TF32): the 3rd launch of each precision is kept.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 2 -c 4 -f -o gpurun_out/r02_halo_dense \
    python tools/ncu_halo.py > gpurun_out/ncu_halo.log 2>&1
tail -3 gpurun_out/ncu_halo.log
