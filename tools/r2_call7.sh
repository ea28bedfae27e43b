#!/bin/bash
set -u
O=gpurun_out/r2_7; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests/test_gpu_dropout_parity.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "greedy" > $O/tests_greedy.log 2>&1
echo "greedy tests exit $?" >> $O/status.txt
timeout 300 python bench.py --steps 30 --warmup 5 --cpu-iters 0 > $O/bench.json 2> $O/bench.err
RECNET_GREEDY_PF=0 timeout 300 python bench.py --steps 30 --warmup 5 --cpu-iters 0 > $O/bench_nogpf.json 2> $O/bench_nogpf.err
cat $O/status.txt; tail -12 $O/tests_greedy.log
python -c "
import json
for n in ('bench','bench_nogpf'):
    d=json.load(open('$O/%s.json' % n)); g=d['workloads']['greedy_b1024']; print(n, d['ms_per_step'], g.get('value'), g.get('ms_per_batch'), g.get('launches_per_batch'), g.get('top_kernel'), g.get('error'))
"
