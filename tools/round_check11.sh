#!/bin/bash
set -u
O=gpurun_out/check11; mkdir -p $O
timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_fq1.json 2> $O/e1
RECNET_FUSED_QUERY=0 timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_fq0.json 2> $O/e2
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/tests_all.log 2>&1
echo "all tests exit $?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/tests_all.log; tail -2 $O/e1
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['launches_per_step'])"; done
