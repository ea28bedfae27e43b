#!/bin/bash
# quick check of the persistent paths: full-size parity, timeline, bench
set -u
O=gpurun_out/r2_q; mkdir -p $O; rm -f $O/*
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "full_size or graph or permutation or linearity or determinism" > $O/tests_fullsize.log 2>&1
echo "fullsize tests exit $?" >> $O/status.txt
timeout 120 python tools/persist_timeline.py > $O/timeline.txt 2>&1
timeout 150 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench.json 2> $O/bench.err
echo "bench exit $?" >> $O/status.txt
cat $O/status.txt; tail -4 $O/tests_fullsize.log
python -c "
import json
d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"
cat $O/timeline.txt
