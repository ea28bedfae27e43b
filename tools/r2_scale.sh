#!/bin/bash
# weak-scaling record on one box: N = 1, 2, 4, 8 back to back (default settings), + the all-reduce check at the largest N
set -u
O=gpurun_out/r2_scale; mkdir -p $O; rm -f $O/*
NMAX=${1:-8}
for N in 1 2 4 8; do
  [ $N -gt $NMAX ] && continue
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_n$N.json 2> $O/bench_n$N.err
  fi
  echo "N=$N exit $?" >> $O/status.txt
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29531 tools/dp_nvlink_check.py > $O/check_n$NMAX.json 2> $O/check_n$NMAX.err
echo "check exit $?" >> $O/status.txt
RECNET_DP_IMPL=nccl RECNET_DP_FLAT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NMAX --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_n${NMAX}_nccl.json 2> $O/bench_n${NMAX}_nccl.err
echo "nccl exit $?" >> $O/status.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29551 tools/step_trace.py --out $O/step_trace_n$NMAX.txt > $O/trace.log 2>&1
echo "trace exit $?" >> $O/status.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/dp_step_check.py > $O/step_check_n2.log 2>&1
echo "step check exit $?" >> $O/status.txt; grep "^{" $O/step_check_n2.log | tail -1
cat $O/status.txt; cat $O/check_n$NMAX.json
for f in $O/bench_*.json; do echo $f; python -c "
import json
try:
    d=json.load(open('$f')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce_impl'), d['clocks'])
except Exception as e: print('bad', e)"; done
