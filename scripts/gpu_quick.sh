#!/bin/bash
N=${N:-2}
mkdir -p /tmp/dpw
SP_DP_PEER=1 SP_STEADY=1 SP_TEST_STEPS=12 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/_dp_gpu_worker.py /tmp/dpw 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -2
python - <<'P'
import numpy as np
s=np.load('/tmp/dpw/single.npz'); r0=np.load('/tmp/dpw/rank0.npz'); r1=np.load('/tmp/dpw/rank1.npz')
print(all(np.array_equal(r0[k], r1[k]) for k in s.files), max(np.max(np.abs(r0[k]-s[k]))/max(np.max(np.abs(s[k])),1e-30) for k in s.files))
P
SP_PEER_TRACE=1 SP_DP_PEER=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/steady_parts.py cnn 128 2>&1 | grep "^cnn\|^loop\|^rank\|Error\|error" | tail -12
