#!/bin/bash
# ncu: launch list of one bench run + full capture of the tcgen05 residual-block convolution.  bash tools/gpu_prof_conv.sh <tag>
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 5 --warmup 3 --cpu-seconds 0.2 > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv2d_umma' -s 40 -c 3 -o $OUT/prof_conv \
  python bench.py --steps 3 --warmup 3 --cpu-seconds 0.2 > $OUT/ncu_full.log 2>&1
ls -la $OUT
