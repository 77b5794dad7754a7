#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "layernorm" 2>&1 | tail -2
python tools/ln_bench.py
