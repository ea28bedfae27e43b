#!/bin/bash
set -u
O=gpurun_out/check6; mkdir -p $O
timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side1.json 2> $O/bench_side1.err
RECNET_SIDE=0 timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side0.json 2> $O/bench_side0.err
timeout 100 python bench.py --steps 200 --warmup 10 --cpu-iters 0 > $O/bench_side1_200.json 2> $O/bench_side1_200.err
tail -2 $O/bench_side1.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['config']['cuda_graph'], d['config'].get('graph_execs'))"; done
