#!/bin/bash
# GPU round with an ncu capture of the transform kernels
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mdct4_fwd|imdct4_inv" --launch-skip 5 --launch-count 13 -f -o gpurun_out/${tag}_mdct python tools/mdct_bench.py --flavours mixed --reps 1 --out gpurun_out/${tag}_mdct_under_ncu.json > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_ncu.log
