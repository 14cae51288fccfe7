#!/bin/bash
# bench + one ncu --set full capture of the kernels named in DESIGN.md (scripts/ncu_targets.py)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-400; tail -5 gpurun_out/bench.err
timeout 300 python scripts/ncu_targets.py > gpurun_out/ncu_plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:"conv_tma2|maxpool|leaky_|adam_multi|colsum|binary_rowvec" -c 16 \
   -o gpurun_out/prof_r01_v2 -f python scripts/ncu_targets.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
