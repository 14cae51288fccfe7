#!/bin/bash
# quick loop: failing-test subset + step profiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_properties.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
for p in tf32 tf32x1; do
  echo "=== step profile cnn $p"
  timeout 300 python scripts/step_profile.py cnn 128 $p 2>&1 | tail -45
done
