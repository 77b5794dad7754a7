#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for mode in "" "--shard"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline $mode > gpurun_out/bench117_n8$mode.log 2>&1; echo "exit $?" >> gpurun_out/bench117_n8$mode.log
python - <<PY
import json
l=[x for x in open('gpurun_out/bench117_n8$mode.log') if x.startswith('{')]
d=json.loads(l[-1]); print(d['n_gpus'], d['config']['parallelism'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
PY
done
