#!/bin/bash
tag=${1:-r02l}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py -x > $O/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest.log
tail -5 $O/${tag}_pytest.log
echo "== new lib"; timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== prev lib"; MDCTGAN_LIB=$PWD/build/variants/lib_prev.so timeout 300 python tools/step_time.py 2>&1 | tail -1
for v in G H; do
  echo "== mdct variant $v"; MDCTGAN_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/mdct_bench.py --flavours mixed --reps 20 --out $O/${tag}_mdct_$v.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: c = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(c['B'], c['T'], 'err', round(c['round_trip_max_err_eps_peak'], 3), {k: (round(v['ms'], 4), round(v['frac_of_hbm_peak'], 3)) for k, v in c.items() if isinstance(v, dict)})
"
done
