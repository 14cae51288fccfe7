#!/bin/bash
# One gpurun call: CUDA-core path tests, tcgen05 bring-up probe, full GPU suite, smoke, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== phase 1: fp32 / CUDA-core path" | tee gpurun_out/phase1.log
SP_PRECISION=fp32 timeout 900 python -m pytest tests -m gpu -q --maxfail=15 \
  -k "not tf32 and not force_umma and not tensor_core" -p no:cacheprovider >> gpurun_out/phase1.log 2>&1
tail -15 gpurun_out/phase1.log
echo "=== phase 2: tcgen05 probe (async producers)" | tee gpurun_out/probe.log
timeout 900 python scripts/umma_probe.py all >> gpurun_out/probe.log 2>&1
grep -E "CASE" gpurun_out/probe.log
echo "=== phase 2b: tcgen05 probe (wait+fence producers)" | tee gpurun_out/probe_sync.log
SP_PROBE_SYNC=1 timeout 900 python scripts/umma_probe.py all >> gpurun_out/probe_sync.log 2>&1
grep -E "CASE" gpurun_out/probe_sync.log
echo "=== phase 3: full gpu suite" | tee gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider >> gpurun_out/pytest_gpu.log 2>&1
tail -40 gpurun_out/pytest_gpu.log
echo "=== phase 4: smoke + bench"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
