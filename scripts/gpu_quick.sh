#!/bin/bash
mkdir -p gpurun_out

timeout 1500 python scripts/op_sweep_grid.py > gpurun_out/r02_op_sweep_grid.log 2>&1; echo rc=$?
tail -5 gpurun_out/r02_op_sweep_grid.log | cut -c1-1500
