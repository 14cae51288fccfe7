#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-100} --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
python - <<P
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','host_issue_ms_per_step')}, 'e2e', d['e2e']['value'], d['regions'], d['steady_replay'])
print(json.dumps(d.get('other_configs'))[:600])
P
tail -5 gpurun_out/bench_n$N.err
