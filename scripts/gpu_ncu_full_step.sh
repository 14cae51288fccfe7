#!/bin/bash
# ncu launch list (warm caches) and ncu --set full of the LAST of 8 eager CIFAR training steps
# (compensated TF32; cudaProfilerStart in scripts/ncu_step.py), and the launch list of the MLP
# step; each only after the same command exited 0 without ncu
mkdir -p gpurun_out
timeout 300 python scripts/ncu_step.py cnn 128 8 > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_step_launches_v3.csv python scripts/ncu_step.py cnn 128 8 > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_mlp_step_launches_v3.csv python scripts/ncu_step.py mlp 200 8 > gpurun_out/ncu_l2.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -o gpurun_out/r02_prof_step_v3 -f python scripts/ncu_step.py cnn 128 8 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/r02_prof_step_v3.ncu-rep gpurun_out/*_v3.csv
