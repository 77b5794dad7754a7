#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainer.py -q -m gpu > gpurun_out/trainer.log 2>&1; echo "exit $?" >> gpurun_out/trainer.log; tail -8 gpurun_out/trainer.log
