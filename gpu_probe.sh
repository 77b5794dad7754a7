#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/attn_bench.py --B 1 --iters 2"
$CMD > gpurun_out/attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 3 -f -o gpurun_out/attn_prof $CMD > gpurun_out/attn_ncu.log 2>&1
echo "exit $?"; cat gpurun_out/attn_plain.log; tail -3 gpurun_out/attn_ncu.log; ls -la gpurun_out/*.ncu-rep
