#!/bin/bash
# GPU round for the train step: parity tests, smoke, bench, ncu launch list of one eager + graphed bench run.  bash tools/gpu_train_prof.sh <tag> [full]
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --cpu-seconds 10 2> $OUT/bench.err | tail -1 > $OUT/bench.json; tail -c 600 $OUT/bench.err; head -c 400 $OUT/bench.json; echo
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
  python tools/train_step_once.py 3 > $OUT/ncu_launches.log 2>&1
tail -3 $OUT/ncu_launches.log
if [ "$2" == "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_wgrad_kernel|conv2d_umma_kernel|conv2d_nhwc_kernel' -s 150 -c 12 -o $OUT/prof_full \
  python tools/train_step_once.py 2 > $OUT/ncu_full.log 2>&1
fi
ls -la $OUT
