#!/bin/bash
# multi-GPU check: DDP parity test + bench line at N GPUs.  bash tools/gpu_multi.sh N tag
N=${1:-2}; tag=${2:-r02n}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q -x > $O/${tag}_ddp_pytest.log 2>&1; tail -3 $O/${tag}_ddp_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-extras > $O/${tag}_bench_${N}gpu.json 2> $O/${tag}_bench_${N}gpu.err; echo "rc=$?"
tail -c 400 $O/${tag}_bench_${N}gpu.err
python - <<PY
import json
for l in open("$O/${tag}_bench_${N}gpu.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "strong", d.get("strong"), "mdct", d["mdct"].get("aggregate"), "longform", {k: d["longform"].get(k) for k in ("seconds_per_clip", "n_gpus")})
PY
