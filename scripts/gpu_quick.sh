#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/steady_parts.py cnn 128 2>&1 | tail -2
timeout 300 python scripts/steady_parts.py vgg 256 2>&1 | tail -2
