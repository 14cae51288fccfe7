#!/bin/bash
mkdir -p gpurun_out
echo "=== probe cases (correctness) im2col"
timeout 600 python scripts/tma2_probe.py all 2>&1 | grep CASE
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
