#!/bin/bash
# full GPU test suite + smoke + one bench line
O=gpurun_out/r2_full; mkdir -p $O; rm -f $O/*
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/tests.log 2>&1
echo "tests exit $?"; tail -5 $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/smoke.log
timeout 300 python bench.py --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"
