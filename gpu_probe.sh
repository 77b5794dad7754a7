#!/bin/bash
# first GPU bring-up: environment probe + GEMM parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
ls /root/reference baseline/_ref > gpurun_out/ls_ref.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu > gpurun_out/gemm_test.log 2>&1
echo "exit $?" >> gpurun_out/gemm_test.log
tail -40 gpurun_out/gemm_test.log
