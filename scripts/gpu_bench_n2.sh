#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','host_issue_ms_per_step')}, d['e2e']['value'], d.get('cuda_graph',{}).get('ms_per_step'))
P
tail -3 gpurun_out/bench_n2.err
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "multi or dist or gpu2 or nccl or parallel" 2>&1 | tail -3
