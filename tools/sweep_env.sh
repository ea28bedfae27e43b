#!/bin/bash
# usage: tools/sweep_env.sh "VAR=a VAR=b ..." -> one short bench per setting, prints samples/s and ms/step
for kv in "$@"; do
  env $kv timeout 200 python bench.py --steps 20 --warmup 3 --cpu-iters 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$kv', d['value'], d['ms_per_step'])"
done
