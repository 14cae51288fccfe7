#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step')}, 'e2e', d['e2e'], 'graph', d['cuda_graph'])
print(d['roofline'])
print(d['cpu_baseline'])
print(d['top_calls_one_step_ms'])
for c in d.get('other_configs',[]): print(json.dumps(c)[:600])
P
tail -3 gpurun_out/bench.err
echo "=== host overhead"
timeout 300 python scripts/host_overhead.py 2>&1 | grep -A16 "host issue time, device idle"
