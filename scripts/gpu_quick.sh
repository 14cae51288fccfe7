#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ncu_step.py cnn 128 8 > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_step_launches_v4.csv python scripts/ncu_step.py cnn 128 8 > gpurun_out/ncu_l.log 2>&1
tail -20 gpurun_out/r02_step_launches_v4.csv | cut -d, -f5,9,15 | cut -c1-150
