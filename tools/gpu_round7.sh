#!/bin/bash
tag=${1:-r02o}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py > $O/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest.log
tail -4 $O/${tag}_pytest.log
timeout 1200 python bench.py --steps 30 --warmup 5 > $O/${tag}_bench.json 2> $O/${tag}_bench.err; echo "bench rc=$?" >> $O/${tag}_bench.err
tail -3 $O/${tag}_bench.err
python - <<PY
import json
d = json.load(open("$O/${tag}_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "reference_api", d["reference_api"]["ms_per_step"], d["reference_api"].get("ms_per_step_all_eager"))
print({k: v.get("ms_per_step") for k, v in d["gpu_baseline"].items()})
PY
