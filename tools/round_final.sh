#!/bin/bash
# Round-end record (one B200): full GPU suite, smoke, the default bench line (with cpu_baseline), the reference arm, and a warm ncu
# launch list of one eager step.  Outputs under gpurun_out/final_g/.
set -u
O=gpurun_out/final_g; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/tests_all.log 2>&1
echo "all tests exit $?" >> $O/status.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/status.txt
timeout 200 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench exit $?" >> $O/status.txt
timeout 200 python bench.py --impl reference --steps 6 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "bench reference exit $?" >> $O/status.txt
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1650 -c 440 --csv --log-file $O/launches_warm.csv python bench.py --steps 1 --warmup 3 --no-graph --cpu-iters 0 > $O/l2.log 2>&1
echo "ncu exit $?" >> $O/status.txt
cat $O/status.txt; tail -2 $O/tests_all.log; tail -2 $O/smoke.log; head -c 600 $O/bench.json; echo; head -c 300 $O/bench_reference.json; echo; wc -l $O/launches_warm.csv
