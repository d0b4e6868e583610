#!/bin/bash
# One gpurun call for a whole round: GPU parity tests, smoke, bench (both arms), ncu launch list.  bash tools/gpu_round.sh [tag]
# (the train-step profile lives in tools/gpu_train_prof.sh; ncu --set full captures in tools/gpu_ncu_full.sh)
TAG=${1:-r01}
bash tools/gpu_train_prof.sh $TAG
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2>/dev/null | tail -1 | tee gpurun_out/$TAG/bench_reference.json | head -c 300; echo
