#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/gpu_tests.log; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench117.log 2>&1; echo "exit $?" >> gpurun_out/bench117.log
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench117.log') if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'], d['kernels_ms_per_step'], d['clocks'])
PY
python tools/hbm_bench.py gpurun_out/hbm_table.md > /dev/null 2>&1
