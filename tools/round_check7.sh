#!/bin/bash
set -u
O=gpurun_out/check7; mkdir -p $O
RECNET_BENCH_ALT=0 timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side1_alt0.json 2> $O/e1
RECNET_BENCH_ALT=1 timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side1_alt1.json 2> $O/e2
RECNET_BENCH_ALT=0 timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side1_alt0b.json 2> $O/e3
RECNET_SIDE=0 RECNET_BENCH_ALT=0 timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side0_alt0.json 2> $O/e4
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config'].get('graph_execs'))"; done
