#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attn_tc" -x > gpurun_out/attn_test.log 2>&1
echo "exit $?" >> gpurun_out/attn_test.log
tail -3 gpurun_out/attn_test.log
timeout 300 python tools/attn_bench.py --B 2 --iters 5 2>&1 | tail -6
