#!/bin/bash
# usage: tools/dp_probe.sh N "ENV=.. ENV=.." ...   -> one short N-rank bench per setting
N=$1; shift
P=29600
for kv in "$@"; do
  P=$((P+1))
  env $kv timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 40 --warmup 5 --cpu-iters 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$kv', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
done
