#!/bin/bash
# full GPU suite, smoke, bench (N=1)
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"
timeout 900 python bench.py --steps ${STEPS:-50} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step','abi_calls_per_step')}, 'e2e', d['e2e']['value'], 'graph', d['cuda_graph'] and d['cuda_graph']['ms_per_step'])
print(d['regions'], d['steady_replay'], d['clocks'])
print(d['roofline'])
print(json.dumps(d.get('other_configs'))[:1200])
print(d['cpu_baseline'])
P
tail -3 gpurun_out/bench.err
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-700
