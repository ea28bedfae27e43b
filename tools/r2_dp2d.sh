#!/bin/bash
# 2-GPU A/B of the data-parallel step under environment switches (same box): bash tools/r2_dp2d.sh "ENV=.." "ENV=.." ...
O=gpurun_out/r2_dp2d; mkdir -p $O; rm -f $O/*
i=0
for cfg in "$@"; do
timeout 600 env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$i bench.py --gpus 2 --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n2_$i.json 2> $O/bench_n2_$i.err
python -c "
import json
d=json.loads([l for l in open('$O/bench_n2_$i.json') if l.startswith('{')][-1]); print('N=2 $cfg', d['value'], d['ms_per_step'], d['e2e']['value'])"
i=$((i+1))
done
