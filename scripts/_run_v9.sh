#!/bin/bash
mkdir -p gpurun_out
echo "=== new tests"
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -p no:cacheprovider -k "skinny or rgb or matmul" 2>&1 | tail -15
echo "=== small-C golden ops"
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "conv or matmul" 2>&1 | tail -5
echo "=== rgb probe"
timeout 300 python scripts/rgb_probe.py 2>&1 | tail -20
echo "=== vgg step"
timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -6
echo "=== bench"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['cuda_graph'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
print(d['top_calls_one_step_ms'])
for c in d.get('other_configs',[]): print(c)
P
tail -3 gpurun_out/bench.err
