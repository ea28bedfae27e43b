#!/bin/bash
# end-of-round measurement pass on one B200: profiles (bench with extras, launch lists, ncu captures) + step timelines
bash tools/collect_profiles_r2.sh > gpurun_out/collect.log 2>&1
tail -3 gpurun_out/collect.log
mkdir -p gpurun_out/trace
timeout 300 python tools/step_trace.py --out gpurun_out/trace/final_lane.txt 2>&1 | tail -1
RECNET_BG_WGRAD=0 RECNET_SIDE=0 RECNET_GEMM_PERSIST=0 timeout 300 python tools/step_trace.py --out gpurun_out/trace/final_r2g_equiv.txt 2>&1 | tail -1
timeout 300 python tools/step_trace.py --recon global --out gpurun_out/trace/final_global.txt 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
tail -c 600 gpurun_out/final/bench_reference.json
