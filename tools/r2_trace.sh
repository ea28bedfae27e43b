#!/bin/bash
mkdir -p gpurun_out/trace
timeout 600 python -m pytest tests/test_gpu_background.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8
RECNET_BG_WGRAD=1 timeout 300 python tools/step_trace.py --out gpurun_out/trace/bg2.txt 2>&1 | tail -1
RECNET_BG_WGRAD=1 RECNET_SIDE=1 timeout 300 python tools/step_trace.py --out gpurun_out/trace/bg2_side.txt 2>&1 | tail -1
