#!/bin/bash
# r02j: new parity tests + MDCT fp64-core tuning sweep + tcgen05 conv phase trace of the cfg4 bottleneck layer + recipe with the TF32 engine
tag=${1:-r02j}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_encodings_gpu.py tests/test_train_gpu.py -m gpu -q -k "encod or reference or round_trip or inference_with or no_lsgan" > $O/${tag}_pytest_new.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest_new.log
tail -25 $O/${tag}_pytest_new.log
for v in A B C D F; do
  lib=build/variants/lib_$v.so; [ $v == A ] && lib=mdctgan_b200/libmdctgan_b200.so
  echo "== variant $v"; MDCTGAN_LIB=$PWD/$lib timeout 300 python tools/mdct_bench.py --flavours mixed --reps 20 --out $O/${tag}_mdct_$v.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: c = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(c['B'], c['T'], 'err', round(c['round_trip_max_err_eps_peak'], 3), {k: (round(v['ms'], 4), round(v['frac_of_hbm_peak'], 3)) for k, v in c.items() if isinstance(v, dict)})
"
done 2>&1 | tee $O/${tag}_mdct_sweep.txt
timeout 300 python tools/umma_trace.py 10 7 > $O/${tag}_umma_trace.txt 2>&1; cat $O/${tag}_umma_trace.txt
timeout 300 python tools/conv_bench.py > $O/${tag}_conv_bench.txt 2>&1; cat $O/${tag}_conv_bench.txt
MDCTGAN_CONV_ENGINE=tf32 timeout 600 python tools/recipe_profile.py 20 > $O/${tag}_recipe_tf32.txt 2>&1; head -30 $O/${tag}_recipe_tf32.txt
