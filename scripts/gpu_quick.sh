#!/bin/bash
for i in 1 2; do
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
done
