#!/bin/bash
for mix in 1 0; do
for ov in 1 0; do
echo "== NCCL_GRAPH_MIXING_SUPPORT=$mix SP_STEADY_DP_OVERLAP=$ov"
NCCL_GRAPH_MIXING_SUPPORT=$mix SP_STEADY_DP_OVERLAP=$ov timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/steady_parts.py cnn 128 2>&1 | grep "^cnn\|^loop\|Error\|error" | tail -4
done
done
