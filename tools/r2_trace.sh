#!/bin/bash
mkdir -p gpurun_out/trace
timeout 300 python tools/step_trace.py --out gpurun_out/trace/final_lane.txt 2>&1 | tail -1
RECNET_BG_WGRAD=0 RECNET_SIDE=0 RECNET_GEMM_PERSIST=0 timeout 300 python tools/step_trace.py --out gpurun_out/trace/final_r2g_equiv.txt 2>&1 | tail -1
timeout 300 python tools/step_trace.py --recon global --out gpurun_out/trace/final_global.txt 2>&1 | tail -1
