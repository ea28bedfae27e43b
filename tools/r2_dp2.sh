#!/bin/bash
# 2-GPU data-parallel variants
set -u
O=gpurun_out/r2_dp2; mkdir -p $O; rm -f $O/*
run() { name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --cpu-iters 0 > $O/$name.json 2> $O/$name.err; echo "$name exit $?" >> $O/status.txt; }
run default A=1
run flat RECNET_DP_FLAT=1
run flat_overlap RECNET_DP_FLAT=1 RECNET_DP_OVERLAP=1
run overlap RECNET_DP_OVERLAP=1
cat $O/status.txt
for f in $O/*.json; do echo $f; python -c "
import json
try:
    d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['allreduce_bytes_per_step'])
except Exception as e: print('bad', e)"; done
