#!/bin/bash
set -u
O=gpurun_out/check2; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_optim.py tests/test_gpu_variants.py -m gpu -q -p no:cacheprovider > $O/tests_new.log 2>&1
echo "new tests exit $?" >> $O/status.txt
RECNET_OPTIMIZER=recnet timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_clipadam.json 2> $O/bench_clipadam.err
echo "bench clipadam exit $?" >> $O/status.txt
timeout 200 python tools/gemm_sweep.py > $O/gemm_sweep.jsonl 2> $O/gemm_sweep.err
echo "sweep exit $?" >> $O/status.txt
cat $O/status.txt; tail -4 $O/tests_new.log; head -c 300 $O/bench_clipadam.json; echo; tail -3 $O/bench_clipadam.err; cut -c1-150 $O/gemm_sweep.jsonl; tail -3 $O/gemm_sweep.err
