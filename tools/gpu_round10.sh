#!/bin/bash
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02zz_bench.json 2> gpurun_out/r02zz_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02zz_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02zz_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"])
print(json.dumps(d["reference_transform_longform"], indent=1)[:3000])
print(d["longform"])
PY
