#!/bin/bash
mkdir -p gpurun_out /tmp/dpw
for st in 0 1; do
echo "== peer path, SP_STEADY=$st"
SP_DP_PEER=1 SP_STEADY=$st SP_TEST_STEPS=12 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/_dp_gpu_worker.py /tmp/dpw 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
python - <<'P'
import numpy as np
s=np.load('/tmp/dpw/single.npz'); r0=np.load('/tmp/dpw/rank0.npz'); r1=np.load('/tmp/dpw/rank1.npz')
print('peer', int(r0['peer_exchange']), 'replayed', int(r0['replayed_updates']))
for k in s.files:
    same=np.array_equal(r0[k], r1[k]); ref=s[k].astype(np.float64)
    err=np.max(np.abs(r0[k]-ref))/max(np.max(np.abs(ref)),1e-30)
    print(k, 'ranks identical', same, 'err vs single', err)
P
done
