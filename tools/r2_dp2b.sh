#!/bin/bash
set -u
O=gpurun_out/r2_dp2b; mkdir -p $O; rm -f $O/*
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dp_nvlink_check.py > $O/check.json 2> $O/check.err
echo "check exit $?" >> $O/status.txt
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --cpu-iters 0 > $O/$name.json 2> $O/$name.err; echo "$name exit $?" >> $O/status.txt; }
run nvlink_overlap A=1
run nvlink_nooverlap RECNET_DP_OVERLAP=0
run nccl_flat RECNET_DP_IMPL=nccl RECNET_DP_FLAT=1
cat $O/status.txt; cat $O/check.json; tail -5 $O/check.err
for f in $O/n*.json; do echo $f; python -c "
import json
try:
    d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['allreduce_bytes_per_step'], d.get('allreduce_impl'))
except Exception as e: print('bad', e)"; tail -3 ${f%.json}.err; done
