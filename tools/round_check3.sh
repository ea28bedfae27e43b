#!/bin/bash
set -u
O=gpurun_out/check3; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_optim.py -m gpu -q -p no:cacheprovider > $O/tests_optim.log 2>&1
echo "optim tests exit $?" >> $O/status.txt
timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_new_plan.json 2> $O/bench_new_plan.err
echo "bench new plan exit $?" >> $O/status.txt
RECNET_GEMM_COSTMODEL=1 timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_old_plan.json 2> $O/bench_old_plan.err
echo "bench old plan exit $?" >> $O/status.txt
RECNET_OPTIMIZER=recnet timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_new_plan_clipadam.json 2> $O/bench_new_plan_clipadam.err
echo "bench new plan + clipadam exit $?" >> $O/status.txt
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -q -p no:cacheprovider -x > $O/tests_all.log 2>&1
echo "all tests exit $?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/tests_optim.log; tail -3 $O/tests_all.log
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"; done
