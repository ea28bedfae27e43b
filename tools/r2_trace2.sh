#!/bin/bash
O=gpurun_out/trace; mkdir -p $O
RECNET_TRACE_ALL_RANKS=1 RECNET_TRACE_ABS=1 RECNET_BG_WGRAD=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 tools/step_trace.py --out $O/dp2abs_lane.txt 2>&1 | tail -2
