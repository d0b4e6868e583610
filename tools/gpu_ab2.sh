#!/bin/bash
for n in 1 2 3 4; do echo "== side streams $n"; MDCTGAN_SIDE_STREAMS=$n timeout 300 python tools/step_time.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -3
