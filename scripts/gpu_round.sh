#!/bin/bash
# full GPU suite, smoke, bench (N=1) and the VGG step time
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"
timeout 900 python bench.py --steps ${STEPS:-200} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step','abi_calls_per_step')}, 'e2e', d['e2e']['value'], 'graph', d['cuda_graph'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])
print(json.dumps(d.get('other_configs'))[:900])
P
tail -3 gpurun_out/bench.err
echo "=== vgg step"
timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -1
