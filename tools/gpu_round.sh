#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + full capture of both kernels.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench ours"; timeout 600 python bench.py 2>&1 | tail -3 | tee $OUT/bench.json
echo "== bench ours fp64"; timeout 600 python bench.py --precision fp64 --cpu-seconds 1 2>&1 | tail -3 | tee $OUT/bench_fp64.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 5 --warmup 3 --cpu-seconds 0.5 > $OUT/ncu_launches.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mdct4_fwd|imdct4_inv' -s 6 -c 2 -o $OUT/prof_full \
  python bench.py --steps 3 --warmup 3 --cpu-seconds 0.5 --e2e-clips 64 > $OUT/ncu_full.log 2>&1
ls -la $OUT
