#!/bin/bash
N=${1:-2}
run() {
  echo "== $1"
  env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --no-extras --no-longform --no-strong --cpu-seconds 0.5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3))
"
}
run "A=0"
run "NCCL_MAX_CTAS=4"
run "NCCL_MAX_CTAS=16 NCCL_MIN_CTAS=16"
run "MDCTGAN_LOW_PRIORITY_STREAMS=sweep_D"
