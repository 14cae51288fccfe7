#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite (PDL on)"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
for pdl in 0 1; do
echo "=== bench SP_PDL=$pdl"
SP_PDL=$pdl timeout 600 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['steps'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'host_issue', round(d['host_issue_ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'graph', d.get('cuda_graph',{}).get('ms_per_step'), d.get('cuda_graph',{}).get('e2e_ms_per_step'))
"
tail -2 gpurun_out/bench.err
done
echo "=== vgg step"
SP_PDL=0 timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -1
SP_PDL=1 timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -1
