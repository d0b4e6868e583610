#!/bin/bash
# ncu --set full of selected kernels of one eager train step.  bash tools/gpu_ncu_full.sh <tag> <kernel regex> <skip> <count>
TAG=${1:-n}; RE=${2:-conv_wgrad_kernel}; SKIP=${3:-60}; CNT=${4:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o $OUT/prof_full -f \
  python tools/train_step_once.py 2 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ncu -i $OUT/prof_full.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
ncu -i $OUT/prof_full.ncu-rep --page source --csv > $OUT/prof_source.csv 2>/dev/null
SZ=$(stat -c %s $OUT/prof_full.ncu-rep); if [ $SZ -gt 30000000 ]; then rm $OUT/prof_full.ncu-rep; fi
ls -la $OUT
