#!/bin/bash
# quick check of a kernel change: parity tests, captured step time (+ optional A/B environment in $1), recipe kernel table head
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py -x 2>&1 | tail -3
timeout 300 python tools/step_time.py 2>&1 | tail -1
if [ -n "$1" ]; then echo "== $1"; env $1 timeout 300 python tools/step_time.py 2>&1 | tail -1; fi
timeout 300 python tools/recipe_profile.py 20 2>&1 | head -6
