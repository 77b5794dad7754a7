#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -q -m gpu > gpurun_out/dp_test.log 2>&1; echo "exit $?" >> gpurun_out/dp_test.log; tail -8 gpurun_out/dp_test.log
