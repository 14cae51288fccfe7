#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
for f in 1 0; do
echo "== SP_FUSE_HEAD=$f"
SP_FUSE_HEAD=$f timeout 300 python scripts/steady_parts.py cnn 128 2>&1 | tail -3
SP_FUSE_HEAD=$f timeout 300 python scripts/steady_parts.py mlp 200 2>&1 | tail -3
done
