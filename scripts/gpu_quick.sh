#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step','abi_calls_per_step','steps','warmup')}, 'e2e', d['e2e'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['traffic'], d['roofline']['bound'])
print(json.dumps(d.get('other_configs'))[:1200])
P
tail -3 gpurun_out/bench.err
