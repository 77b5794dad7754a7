#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_curve_gpu.py -q -s -m gpu > gpurun_out/curve.log 2>&1; echo "exit $?" >> gpurun_out/curve.log; tail -12 gpurun_out/curve.log
