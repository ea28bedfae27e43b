#!/bin/bash
set -u
O=gpurun_out/r2_6; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropout_parity.py -m gpu -q -p no:cacheprovider -x -k "full_size or graph or permutation" > $O/tests_fullsize.log 2>&1
echo "fullsize tests exit $?" >> $O/status.txt
timeout 300 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench.json 2> $O/bench.err
echo "bench exit $?" >> $O/status.txt
RECNET_PERSIST_GLOBAL=0 timeout 300 python bench.py --steps 20 --warmup 5 --cpu-iters 0 --recon global --no-extras > $O/bench_global_old.json 2> $O/bench_global_old.err
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-iters 0 --recon global --no-extras > $O/bench_global_new.json 2> $O/bench_global_new.err
cat $O/status.txt; tail -4 $O/tests_fullsize.log
python -c "
import json
d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('ms_per_step', v.get('ms_per_batch')), v.get('launches_per_step'), v.get('error'))
for n in ('old','new'):
    g=json.load(open('$O/bench_global_%s.json' % n)); print('global', n, g['value'], g['ms_per_step'], g['launches_per_step'])
"
