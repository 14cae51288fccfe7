#!/bin/bash
# ncu --set full of one eager CIFAR training step (21 kernels, compensated TF32); only after the
# same command exited 0 without ncu
mkdir -p gpurun_out
timeout 300 python scripts/ncu_step.py cnn 128 8 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -s 126 -c 21 \
   -o gpurun_out/r02_prof_step -f python scripts/ncu_step.py cnn 128 8 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/r02_prof_step.ncu-rep
