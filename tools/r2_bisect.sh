#!/bin/bash
for cfg in "RECNET_SIDE=0 RECNET_GEMM_PERSIST=1" "RECNET_SIDE=1 RECNET_GEMM_PERSIST=0" "RECNET_SIDE=0 RECNET_GEMM_PERSIST=0" "RECNET_SIDE=1 RECNET_GEMM_PERSIST=1"; do
  echo "== $cfg"
  env $cfg timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "shape_variants" 2>&1 | grep -E "passed|failed|AssertionError:" | head -5
done
