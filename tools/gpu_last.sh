#!/bin/bash
# last sanity round of r02: parity tests, smoke, bench line
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py > $O/r02_last_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_last_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 30 --warmup 5 > $O/r02_last_bench.json 2> $O/r02_last_bench.err; echo "bench rc=$?"; tail -2 $O/r02_last_bench.err
python - <<PY
import json
d = json.load(open("$O/r02_last_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "longform", d["longform"].get("seconds_per_clip"), d["longform"].get("seconds_per_clip_gen_overlap_0"))
PY
