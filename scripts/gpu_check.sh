#!/bin/bash
# quick: probe + suite + smoke + bench + op sweep (no ncu)
mkdir -p gpurun_out
echo "=== tcgen05 probe"; timeout 900 python scripts/umma_probe.py all > gpurun_out/probe.log 2>&1; grep -E "CASE" gpurun_out/probe.log
bash scripts/gpu_profile.sh
