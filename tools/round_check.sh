#!/bin/bash
# One-call round check (run under gpurun on one B200, ~1.5 min): whole GPU suite, smoke(), and the bench line with an optional
# A/B switch:  bash tools/round_check.sh [ENV=VALUE ...]   e.g.  bash tools/round_check.sh RECNET_SIDE=1
# Everything lands under gpurun_out/check/.
set -u
O=gpurun_out/check; mkdir -p $O; rm -f $O/status.txt
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/tests_all.log 2>&1
echo "gpu tests exit $?" >> $O/status.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/status.txt
timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_default.json 2> $O/bench_default.err
echo "bench default exit $?" >> $O/status.txt
if [ $# -gt 0 ]; then
  env "$@" timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_variant.json 2> $O/bench_variant.err
  echo "bench variant ($*) exit $?" >> $O/status.txt
fi
cat $O/status.txt; tail -3 $O/tests_all.log; tail -2 $O/smoke.log
for f in $O/bench_*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['launches_per_step'])"; done
