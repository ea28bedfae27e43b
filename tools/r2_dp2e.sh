#!/bin/bash
# 2-GPU end check: whole data-parallel steps vs NCCL (eager + graph), then one bench line
O=gpurun_out/r2_dp2e; mkdir -p $O; rm -f $O/*
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_step_check.py > $O/step_check.log 2>&1
echo "step check exit $?"; grep "^{" $O/step_check.log | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n2.json 2> $O/bench_n2.err
python -c "
import json
d=json.loads([l for l in open('$O/bench_n2.json') if l.startswith('{')][-1]); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'])"
