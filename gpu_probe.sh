#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attn_tc" -x 2>&1 | tail -2
timeout 120 python tools/attn_bench.py --B 2 --iters 5 2>&1 | tail -6
