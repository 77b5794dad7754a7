#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x > gpurun_out/gemm_test.log 2>&1; echo "exit $?" >> gpurun_out/gemm_test.log; tail -5 gpurun_out/gemm_test.log
timeout 200 python tools/gemm_bench.py 2>&1 | tail -13
