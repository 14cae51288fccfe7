#!/bin/bash
# bench (with extras) + ncu full capture of the conv kernels of a short bench run
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/ncu_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:"conv_tma2|conv_wgrad_rgb|conv_fprop_rgb|igemm_umma" -s 40 -c 10 \
   -o gpurun_out/prof_bench_v3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log | cut -c1-200
