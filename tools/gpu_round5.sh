#!/bin/bash
# parity tests + A/B of the split-K cluster size + conv layer bench + bench line
tag=${1:-r02k}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py > $O/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest.log
tail -12 $O/${tag}_pytest.log
echo "== split cap 16"; timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== split cap 8"; MDCTGAN_UMMA_MAX_SPLIT=8 timeout 300 python tools/step_time.py 2>&1 | tail -1
timeout 300 python tools/umma_trace.py 10 > $O/${tag}_umma_trace.txt 2>&1; head -14 $O/${tag}_umma_trace.txt
timeout 300 python tools/conv_bench.py > $O/${tag}_conv_bench.txt 2>&1; cat $O/${tag}_conv_bench.txt
timeout 300 python tools/mdct_bench.py --flavours mixed --reps 20 --out $O/${tag}_mdct.json 2>&1 | cut -c1-400
if [ "$2" == "bench" ]; then
timeout 1200 python bench.py --steps 30 --warmup 5 > $O/${tag}_bench.json 2> $O/${tag}_bench.err; echo "bench rc=$?" >> $O/${tag}_bench.err
tail -3 $O/${tag}_bench.err
fi
