#!/bin/bash
python scripts/steady_parts.py mlp 200 2>&1 | tail -2
SP_STEADY=0 python scripts/step_profile.py mlp 200 2>&1 | tail -30
