#!/bin/bash
# Refresh committed evidence: bench JSON, launch list, ncu full captures (tcgen05 conv at VGG
# sizes + the HBM-bound kernels), op sweep.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 300 gpurun_out/bench.json; echo
timeout 900 python scripts/op_sweep.py --reps 10 > gpurun_out/op_sweep.log 2>&1; tail -3 gpurun_out/op_sweep.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 260 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 300 python scripts/ncu_targets.py > gpurun_out/ncu_plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"igemm_umma|maxpool|leaky|adam_multi|colsum_vec4" -c 14 \
   -o gpurun_out/prof_r01 -f python scripts/ncu_targets.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
