#!/bin/bash
O=gpurun_out/r2_dp2d; mkdir -p $O; rm -f $O/*
i=0
for cfg in "RECNET_BG_WGRAD=1" "RECNET_BG_WGRAD=0" "RECNET_BG_WGRAD=1 RECNET_BG_CTAS=32"; do
timeout 600 env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$i bench.py --gpus 2 --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n2_$i.json 2> $O/bench_n2_$i.err
python -c "
import json
d=json.loads([l for l in open('$O/bench_n2_$i.json') if l.startswith('{')][-1]); print('N=2 $cfg', d['value'], d['ms_per_step'], d['e2e']['value'])"
i=$((i+1))
done
for cfg in "RECNET_BG_WGRAD=1" "RECNET_BG_WGRAD=1 RECNET_BG_CTAS=32"; do
env $cfg timeout 300 python bench.py --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.load(open('$O/bench_n1.json')); print('N=1 $cfg', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
