#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/gpu_tests.log; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench117.log 2>&1; echo "exit $?" >> gpurun_out/bench117.log
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench117.log') if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['kernels_ms_per_step'], d['clocks'], d['cpu_baseline']['value'])
PY
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
