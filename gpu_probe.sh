#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu > gpurun_out/kern_test.log 2>&1
echo "exit $?" >> gpurun_out/kern_test.log
tail -60 gpurun_out/kern_test.log
