#!/bin/bash
# One gpurun call: full GPU suite, bench, host-overhead profile, then the ncu launch list of a
# short bench run (only after the same command exited 0 without ncu).
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
echo "=== bench"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== host overhead"
timeout 300 python scripts/host_overhead.py > gpurun_out/host_overhead.log 2>&1; head -30 gpurun_out/host_overhead.log
echo "=== ncu launch list"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
