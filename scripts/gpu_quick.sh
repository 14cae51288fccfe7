#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
python scripts/e2e_phases.py cnn 2>&1 | tail -2
SP_EAGER_UPLOAD=0 python scripts/e2e_phases.py cnn 2>&1 | tail -2
python scripts/e2e_phases.py mlp 2>&1 | tail -2
