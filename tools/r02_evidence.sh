set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_final.log
tail -3 gpurun_out/r02_gputest_final.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_launch_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc|attn_bwd_fused" -c 2 -o gpurun_out/r02_attn_final python tools/attn_bench.py --B 8 --iters 1 > gpurun_out/r02_ncu_attn_final.log 2>&1
ncu -i gpurun_out/r02_attn_final.ncu-rep --page raw --csv > gpurun_out/r02_attn_final_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:gemm_tc -c 3 -o gpurun_out/r02_gemm_pair python tools/gemm_bench.py "qkv  fwd" > gpurun_out/r02_gemm_pair_ncu.log 2>&1
ncu -i gpurun_out/r02_gemm_pair.ncu-rep --page raw --csv > gpurun_out/r02_gemm_pair_raw.csv 2>/dev/null
python tools/gemm_bench.py > gpurun_out/r02_gemm_bench_final.log 2>&1
ls -la gpurun_out | tail -12
