#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm_all_operand" > gpurun_out/gemm_tests.log 2>&1
tail -12 gpurun_out/gemm_tests.log
timeout 900 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_r2.jsonl 2> gpurun_out/gemm_sweep_r2.err
tail -3 gpurun_out/gemm_sweep_r2.err
python - <<'PY'
import json
for l in open("gpurun_out/gemm_sweep_r2.jsonl"):
    d=json.loads(l); u=d["us"]
    old=min(v for k,v in u.items() if not k.startswith("p") and isinstance(v,float))
    print(d["shape"], d["MNK"], "old", old, {k:v for k,v in u.items() if k.startswith("p")})
PY
