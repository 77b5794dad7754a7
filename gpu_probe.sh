#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiles_gpu.py -q -m gpu > gpurun_out/tiles.log 2>&1; echo "exit $?" >> gpurun_out/tiles.log; tail -5 gpurun_out/tiles.log
