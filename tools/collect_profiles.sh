#!/bin/bash
# Round-end measurement pass (run under gpurun on one B200): bench line, ncu launch lists (cold = recipe default, warm = L2 as
# the previous kernel left it), one `ncu --set full` capture per hot-loop kernel.  Outputs under gpurun_out/final/.
set -u
O=gpurun_out/final; mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --no-graph --cpu-iters 0"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv &
SMI=$!
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err
kill $SMI
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1650 -c 440 --csv --log-file $O/launches_cold.csv $CMD > $O/l1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1650 -c 440 --csv --log-file $O/launches_warm.csv $CMD > $O/l2.log 2>&1
for k in pf_fwd_kernel pf_bwd_kernel lean_fwd_kernel lean_bwd_kernel lstm_cell_fwd_kernel lstm_cell_bwd_kernel adam_mt_kernel stage_multi_kernel; do
  timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$k -s 12 -c 1 -f -o $O/$k $CMD > $O/$k.log 2>&1
done
# the per-step gate GEMM of the local reconstructor ([100 x 2048] x [2048 x 6144]): warm and cold
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<\(int\)128, \(int\)3, \(bool\)0, \(bool\)0>" -s 12 -c 1 -f -o $O/gate_gemm_warm $CMD > $O/g1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<\(int\)128, \(int\)3, \(bool\)0, \(bool\)0>" -s 12 -c 1 -f -o $O/gate_gemm_cold $CMD > $O/g2.log 2>&1
# the largest batched GEMM (local reconstructor dW_hh, 6144 x 1536 x 2800, 256-wide tiles)
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<\(int\)256" -s 8 -c 1 -f -o $O/wgrad_gemm $CMD > $O/g3.log 2>&1
# keep what travels back small: csv pages instead of the 15 MB reports (gpurun merges at most 64 MiB)
for r in $O/*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $r --page source --csv > ${b}_source.csv 2>/dev/null
  rm -f $r
done
ls -la $O
