#!/bin/bash
set -u
O=gpurun_out/r2_4; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests/test_gpu_dropout_parity.py -m gpu -q -p no:cacheprovider > $O/tests_dropout.log 2>&1
echo "dropout tests exit $?" >> $O/status.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_dropout_parity.py > $O/tests_all.log 2>&1
echo "all tests exit $?" >> $O/status.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/status.txt
timeout 150 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench.json 2> $O/bench.err
echo "bench exit $?" >> $O/status.txt
cat $O/status.txt; tail -25 $O/tests_dropout.log; tail -5 $O/tests_all.log; tail -3 $O/smoke.log
python -c "
import json
d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"
