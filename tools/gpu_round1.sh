#!/bin/bash
# First GPU pass: op parity, tcgen05 bring-up, golden vectors, reference timing.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_ops.log
timeout 300 python -m pytest tests/test_igemm_gpu.py -q -m gpu -x 2>&1 | tail -60 > gpurun_out/t_igemm.log
timeout 300 python -m pytest tests/test_pvconv_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_pvconv.log
timeout 300 python tests/golden/make_golden.py gpurun_out/ref_ops_golden.npz > gpurun_out/golden.log 2>&1
timeout 300 python tests/tools/ref_gpu_time.py > gpurun_out/ref_time.log 2>&1
tail -5 gpurun_out/t_ops.log gpurun_out/t_igemm.log gpurun_out/t_pvconv.log gpurun_out/golden.log gpurun_out/ref_time.log
