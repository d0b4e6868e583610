#!/bin/bash
# A/B of library builds on the captured train step: bash tools/gpu_ab.sh
for rep in 1 2; do
echo "== prev lib"; MDCTGAN_LIB=$PWD/build/variants/lib_prev.so timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== new lib, split cap 16"; timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== new lib, split cap 8"; MDCTGAN_UMMA_MAX_SPLIT=8 timeout 300 python tools/step_time.py 2>&1 | tail -1
done
timeout 300 python tools/conv_bench.py 2>&1 | cut -c1-110
