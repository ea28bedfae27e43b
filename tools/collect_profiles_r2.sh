#!/bin/bash
# Round-2 measurement pass (one B200): full bench line, warm ncu launch list of one eager step, `ncu --set full` captures of the
# persistent loop kernels, the cluster decoder loop, the fused decoder backward kernel and the largest batched GEMM.
# Outputs under gpurun_out/final/ (csv pages instead of .ncu-rep files: gpurun merges at most 64 MiB).
set -u
O=gpurun_out/final; mkdir -p $O; rm -f $O/*
CMD="python bench.py --steps 1 --warmup 3 --no-graph --cpu-iters 0 --no-extras"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv &
SMI=$!
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err
kill $SMI
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 560 -c 150 --csv --log-file $O/launches_warm.csv $CMD > $O/l2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 150 --csv --log-file $O/launches_cold.csv $CMD > $O/l1.log 2>&1
for k in local_fwd_kernel local_bwd_kernel decoder_fwd_cluster_kernel pf_bwd_kernel adam_mt_kernel; do
  timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -f -o $O/$k $CMD > $O/$k.log 2>&1
done
# persistent batched GEMMs (gemm_tc2.cuh): the largest weight gradient (CTA pairs) and the vocabulary projection (2nd <256,4,K-major> launch of a step)
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc2_kernel<\(int\)256, \(int\)6" -s 3 -c 1 -f -o $O/wgrad_gemm_pair $CMD > $O/g3.log 2>&1
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc2_kernel<\(int\)256, \(int\)4, \(bool\)0, \(bool\)0, \(int\)1, \(bool\)0>" -s 10 -c 1 -f -o $O/logits_gemm $CMD > $O/g5.log 2>&1
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<\(int\)64, \(int\)4, \(bool\)0, \(bool\)1>" -s 12 -c 1 -f -o $O/dec_dh_gemm $CMD > $O/g4.log 2>&1
for r in $O/*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $r --page source --csv > ${b}_source.csv 2>/dev/null
  rm -f $r
done
ls -la $O
