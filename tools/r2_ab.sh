#!/bin/bash
# same-box A/B of env switches on the headline bench: usage r2_ab.sh "ENV1=.. ENV2=.." "ENV..." ...
O=gpurun_out/r2_ab; mkdir -p $O; rm -f $O/*
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "full_size or graph or permutation or linearity or determinism" > $O/tests_fullsize.log 2>&1
echo "fullsize tests exit $?"; tail -3 $O/tests_fullsize.log
i=0
for cfg in "$@"; do
  for rep in 1 2; do
    env $cfg timeout 200 python bench.py --steps 50 --warmup 5 --cpu-iters 0 --no-extras > $O/bench_${i}_$rep.json 2> $O/bench_${i}_$rep.err
    python -c "
import json
d=json.load(open('$O/bench_${i}_$rep.json')); print('$cfg', d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"
  done
  i=$((i+1))
done
