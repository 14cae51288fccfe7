#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12
timeout 600 python scripts/op_sweep_grid.py --quick --no-numpy 2>&1 | grep "s2:" | head -20
