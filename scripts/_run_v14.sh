#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
echo "=== ctypes fallback subset"
SP_NO_FASTCALL=1 timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "=== host overhead"
timeout 300 python scripts/host_overhead.py 2>&1 | grep -A16 "host issue time, device idle"
echo "=== bench"
for i in 1 2; do
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'graph', d['cuda_graph']['value'], d['cuda_graph']['e2e_value'], d['roofline']['kernel'], d['roofline']['frac'])
for c in d.get('other_configs',[]): print(json.dumps(c)[:400])
P
done
tail -3 gpurun_out/bench.err
