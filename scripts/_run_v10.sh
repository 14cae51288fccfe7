#!/bin/bash
mkdir -p gpurun_out
echo "=== new tests"
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -p no:cacheprovider -k "skinny or rgb or matmul" 2>&1 | tail -15
echo "=== golden ops"
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "conv or matmul" 2>&1 | tail -5
echo "=== rgb probe"
timeout 300 python scripts/rgb_probe.py 2>&1 | grep tf32 | cut -c1-330
for r in 2 3 5 6 10; do echo "--- wgrad rows $r"; SP_RGB_WGRAD_ROWS=$r SP_PROBE_ONLY=cifar_l1,vgg_l1 timeout 300 python scripts/rgb_probe.py 2>&1 | grep tf32 | grep wgrad | cut -c1-200; done
echo "=== vgg step"
timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -3
echo "=== bench"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['cuda_graph'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
print(d['top_calls_one_step_ms'])
for c in d.get('other_configs',[]): print(c)
P
tail -3 gpurun_out/bench.err
echo "=== models tests"
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_reference_suite.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
