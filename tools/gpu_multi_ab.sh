#!/bin/bash
N=${1:-4}
timeout 300 python tools/step_time.py 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-extras --no-longform --no-strong --cpu-seconds 0.5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
"
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q -x 2>&1 | tail -2
