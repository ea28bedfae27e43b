#!/bin/bash
# r2 call 1: experimental (never-run) paths, fresh baseline, segment timing
set -u
O=gpurun_out/r2_1; mkdir -p $O; rm -f $O/status.txt
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
RECNET_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_zz_optin.py -m gpu -q -p no:cacheprovider -k experimental > $O/tests_experimental.log 2>&1
echo "experimental tests exit $?" >> $O/status.txt
timeout 150 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_default.json 2> $O/bench_default.err
echo "bench default exit $?" >> $O/status.txt
RECNET_DEC_CLUSTER=1 timeout 150 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_deccluster.json 2> $O/bench_deccluster.err
echo "bench dec cluster exit $?" >> $O/status.txt
RECNET_DEFER_REG=1 timeout 150 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_deferreg.json 2> $O/bench_deferreg.err
echo "bench defer reg exit $?" >> $O/status.txt
timeout 200 python tools/segments.py > $O/segments_local.json 2> $O/segments_local.err
echo "segments exit $?" >> $O/status.txt
timeout 200 python tools/segments.py --recon none > $O/segments_none.json 2> $O/segments_none.err
cat $O/status.txt; tail -15 $O/tests_experimental.log
for f in $O/bench_*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"; done
cat $O/segments_local.json $O/segments_none.json
