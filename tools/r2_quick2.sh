#!/bin/bash
O=gpurun_out/r2_q2; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_background.py -m gpu -q -p no:cacheprovider -x -k "full_size or graph or golden or shape_variants or background" > $O/tests.log 2>&1
echo "tests exit $?"; tail -3 $O/tests.log
for cfg in "$@"; do
  for rep in 1 2; do
    env $cfg timeout 200 python bench.py --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench.json 2> $O/bench.err
    python -c "
import json
d=json.load(open('$O/bench.json')); print('$cfg', d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"
  done
done
timeout 300 python tools/step_trace.py --out gpurun_out/trace/q2.txt 2>&1 | tail -1
