#!/bin/bash
O=gpurun_out/r2_dp2c; mkdir -p $O; rm -f $O/*
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_step_check.py > $O/step_check.log 2>&1
echo "step check exit $?"; grep "^{" $O/step_check.log | tail -1; tail -3 $O/step_check.log | cut -c1-300
for rep in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$rep bench.py --gpus 2 --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n2_$rep.json 2> $O/bench_n2_$rep.err
echo "bench exit $?"
python -c "
import json
d=json.loads([l for l in open('$O/bench_n2_$rep.json') if l.startswith('{')][-1]); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 300 python bench.py --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.load(open('$O/bench_n1.json')); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'])"
