#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/attn_bench.py --B 8 --iters 1"
$CMD > gpurun_out/attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 9 -c 3 -f -o gpurun_out/attn_prof_final $CMD > gpurun_out/attn_ncu.log 2>&1
echo "exit $?"; cat gpurun_out/attn_plain.log; tail -2 gpurun_out/attn_ncu.log
