#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_properties.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "=== vgg step (fusion on / conv-leaky fusion off)"
timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -1
SP_FUSE_CONV_LEAKY=0 timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -1
echo "=== dgrad timing"
SP_PROBE_B2B=1 SP_PROBE_ONLY=vgg_conv1b_64_64,vgg_conv2b_128_128,vgg_conv3b_256_256 timeout 300 python scripts/tma2_probe.py time 2>&1 | grep dgrad | cut -c1-200
SP_PROBE_HALO=1 SP_PROBE_B2B=1 SP_PROBE_ONLY=vgg_conv1b_64_64,vgg_conv2b_128_128 timeout 300 python scripts/tma2_probe.py time 2>&1 | grep dgrad | cut -c1-200
