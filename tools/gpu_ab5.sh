#!/bin/bash
for v in U5 U4 U12; do
  echo "== mdct variant $v"; MDCTGAN_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/mdct_bench.py --flavours mixed --reps 20 --out gpurun_out/mdct_$v.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: c = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(c['B'], c['T'], 'err', round(c['round_trip_max_err_eps_peak'], 3), round(c['round_trip_mean_err_eps_peak'], 3), {k: (round(v['ms'], 4), round(v['frac_of_hbm_peak'], 3)) for k, v in c.items() if isinstance(v, dict)})
"
done
