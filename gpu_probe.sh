#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -q -m gpu > gpurun_out/dp_test.log 2>&1; echo "exit $?" >> gpurun_out/dp_test.log; tail -6 gpurun_out/dp_test.log
timeout 300 python bench.py --workload 8m --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('8m:', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
