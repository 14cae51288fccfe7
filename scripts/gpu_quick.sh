#!/bin/bash
for bn in 0 64; do
echo "== SP_TMA2_BN=$bn"
SP_TMA2_BN=$bn python scripts/steady_parts.py cnn 128 2>&1 | tail -2
done
