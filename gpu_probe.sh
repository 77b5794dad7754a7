#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attn_tc" -x 2>&1 | tail -4
timeout 120 python tools/attn_bench.py --B 2 --iters 5 --fwd-only 2>&1 | tail -1
