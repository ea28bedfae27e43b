#!/bin/bash
# One-call round check (run under gpurun on one B200): new GPU tests first, then the whole GPU suite, smoke(), and the bench line
# with both optimiser implementations.  Everything lands under gpurun_out/check/.
set -u
O=gpurun_out/check; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
K="gru or single_step or per_step or two_layer or beam or greedy or cuda_graph or checkpoint"
timeout 200 python -m pytest tests/test_gpu_optim.py tests/test_gpu_variants.py -m gpu -q -s -p no:cacheprovider > $O/tests_new.log 2>&1
echo "new tests exit $?" >> $O/status.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "$K" > $O/tests_touched.log 2>&1
echo "touched tests exit $?" >> $O/status.txt
timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 2 > $O/bench_torch_adam.json 2> $O/bench_torch_adam.err
echo "bench torch exit $?" >> $O/status.txt
RECNET_OPTIMIZER=recnet timeout 100 python bench.py --steps 50 --warmup 5 --cpu-iters 0 > $O/bench_clipadam.json 2> $O/bench_clipadam.err
echo "bench clipadam exit $?" >> $O/status.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/status.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "not ($K)" --durations=10 > $O/tests_all.log 2>&1
echo "remaining tests exit $?" >> $O/status.txt
cat $O/status.txt; tail -5 $O/tests_new.log; tail -3 $O/tests_touched.log; tail -3 $O/tests_all.log; cat $O/bench_torch_adam.json | head -c 400; echo; cat $O/bench_clipadam.json | head -c 400
