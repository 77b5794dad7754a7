#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -k "frontend or golden or 8m" 2>&1 | tail -3
python tools/hbm_bench.py 2>&1 | grep "front end"
