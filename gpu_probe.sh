#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attn_tc" -x > gpurun_out/attn_test.log 2>&1
echo "exit $?" >> gpurun_out/attn_test.log
tail -40 gpurun_out/attn_test.log
