#!/bin/bash
# new tests, then the ncu launch list of a short bench run (only after the same command exited 0
# without ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 120 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
tail -30 gpurun_out/launches.csv | cut -c1-160
