#!/bin/bash
set -u
O=gpurun_out/check5; mkdir -p $O
timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side1.json 2> $O/bench_side1.err
echo "bench side=1 exit $?" >> $O/status.txt
RECNET_SIDE=0 timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side0.json 2> $O/bench_side0.err
echo "bench side=0 exit $?" >> $O/status.txt
timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_side1b.json 2> $O/bench_side1b.err
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/tests_all.log 2>&1
echo "all tests exit $?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/tests_all.log; tail -2 $O/bench_side1.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['config']['cuda_graph'])"; done
