#!/bin/bash
set -u
O=gpurun_out/r2_3; mkdir -p $O; rm -f $O/*
timeout 120 python tools/persist_timeline_bwd.py > $O/timeline_bwd.txt 2>&1
timeout 200 python tools/segments.py > $O/segments_local.json 2> $O/segments_local.err
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/tests_all.log 2>&1
echo "all tests exit $?" >> $O/status.txt
cat $O/status.txt; tail -5 $O/tests_all.log; cat $O/segments_local.json; cat $O/timeline_bwd.txt
