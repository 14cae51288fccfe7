#!/bin/bash
# scratch: whatever the current experiment needs (default: the full GPU suite)
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
