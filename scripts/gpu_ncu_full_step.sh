#!/bin/bash
# ncu --set full of one training step (24 kernels) of a short bench run; only after the same
# command exited 0 without ncu
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -s 779 -c 24 \
   -o gpurun_out/prof_step_v4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/ncu_full4.log 2>&1
tail -2 gpurun_out/ncu_full4.log | cut -c1-200
ls -la gpurun_out/prof_step_v4.ncu-rep
