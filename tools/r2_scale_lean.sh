#!/bin/bash
# lean weak-scaling record on one 8-GPU box: N = 1, 4, 8 back to back (N = 2 comes from tools/r2_dp2d.sh on a 2-GPU box)
O=gpurun_out/r2_scale2; mkdir -p $O; rm -f $O/*
timeout 200 python bench.py --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
for N in 8 4; do
  timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n$N.json 2> $O/bench_n$N.err
  echo "N=$N exit $?"
done
for f in $O/bench_n1.json $O/bench_n4.json $O/bench_n8.json; do python -c "
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce_impl'), d['clocks'])
except Exception as e: print('bad', e)"; done
