#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py -x 2>&1 | tail -4
for f in 1 0 1 0; do echo "== fold fusion $f"; MDCTGAN_FOLD_FUSION=$f timeout 300 python tools/step_time.py 2>&1 | tail -1; done
