#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/op_sweep_grid.py --out gpurun_out/r02_op_sweep_grid_v2.json > gpurun_out/r02_op_sweep_grid_v2.log 2>&1; echo sweep rc=$?
tail -2 gpurun_out/r02_op_sweep_grid_v2.log | cut -c1-600
bash scripts/gpu_round.sh
