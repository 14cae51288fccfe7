#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
SP_STEADY_SPECULATE=1 timeout 600 python -m pytest tests/test_gpu_steady.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
