#!/bin/bash
set -u
O=gpurun_out/r2_2; mkdir -p $O; rm -f $O/status.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "full_size or graph or permutation or linearity" > $O/tests_fullsize.log 2>&1
echo "fullsize tests exit $?" >> $O/status.txt
timeout 150 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_persist.json 2> $O/bench_persist.err
echo "bench persist exit $?" >> $O/status.txt
RECNET_PERSIST=0 timeout 150 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_nopersist.json 2> $O/bench_nopersist.err
echo "bench nopersist exit $?" >> $O/status.txt
timeout 120 python tools/persist_timeline.py > $O/timeline.txt 2>&1
echo "timeline exit $?" >> $O/status.txt
timeout 200 python tools/segments.py > $O/segments_local.json 2> $O/segments_local.err
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/tests_all.log 2>&1
echo "all tests exit $?" >> $O/status.txt
cat $O/status.txt; tail -15 $O/tests_fullsize.log; tail -5 $O/tests_all.log
for f in $O/bench_*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"; done
cat $O/segments_local.json; cat $O/timeline.txt
