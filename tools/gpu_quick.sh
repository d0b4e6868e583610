#!/bin/bash
# Quick GPU iteration: parity tests + bench (+ optional ncu full capture).  bash tools/gpu_quick.sh <tag> [ncu]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench ours"; timeout 600 python bench.py --cpu-seconds 2 2>&1 | tail -3 | tee $OUT/bench.json
if [ "$2" == "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mdct4_fwd|imdct4_inv' -s 6 -c 2 -o $OUT/prof_full \
  python bench.py --steps 3 --warmup 3 --cpu-seconds 0.5 --e2e-clips 64 > $OUT/ncu_full.log 2>&1
fi
