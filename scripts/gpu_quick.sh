#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/steady_parts.py mlp 200 2>&1 | tail -1
timeout 300 python scripts/steady_parts.py cnn 128 2>&1 | tail -1
