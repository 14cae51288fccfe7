#!/bin/bash
# ncu --set full of the 17 conv kernels of one VGG-6 (64x64, batch 256) training step; only after
# the same command exited 0 without ncu
mkdir -p gpurun_out
timeout 300 python scripts/vgg_step.py 4 > gpurun_out/vgg_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_" -s 34 -c 17 \
   -o gpurun_out/prof_vgg_convs -f python scripts/vgg_step.py 4 > gpurun_out/ncu_vgg_full.log 2>&1
tail -2 gpurun_out/ncu_vgg_full.log | cut -c1-200
ls -la gpurun_out/prof_vgg_convs.ncu-rep
