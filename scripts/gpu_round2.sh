#!/bin/bash
# One gpurun call: 2-SM TMA probe, full GPU suite, smoke, bench, op sweep.
mkdir -p gpurun_out
echo "=== tma2 probe"; timeout 600 python scripts/tma2_probe.py all > gpurun_out/tma2_probe.log 2>&1; grep -c PASS gpurun_out/tma2_probe.log; grep -E "FAIL|CRASH|TIMEOUT" gpurun_out/tma2_probe.log
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
echo "=== smoke + bench"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== op sweep"
timeout 600 python scripts/op_sweep.py > gpurun_out/op_sweep.log 2>&1; tail -60 gpurun_out/op_sweep.log | cut -c1-200
