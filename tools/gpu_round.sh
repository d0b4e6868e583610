#!/bin/bash
# One GPU round: parity tests, the bench line.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r02a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 ${BENCH_EXTRA} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
grep -E "passed|failed" gpurun_out/${tag}_pytest.log | tail -2; grep -E "^FAILED" gpurun_out/${tag}_pytest.log | head -20; tail -3 gpurun_out/${tag}_bench.err
