#!/bin/bash
echo "== PDL on"; timeout 300 python tools/conv_bench.py 2>&1 | cut -c1-100
echo "== PDL off"; MDCTGAN_PDL=0 timeout 300 python tools/conv_bench.py 2>&1 | cut -c1-100
for p in 1 0 1 0; do echo "== step PDL $p"; MDCTGAN_PDL=$p timeout 300 python tools/step_time.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_nn_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -3
