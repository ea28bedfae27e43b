#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropout_parity.py tests/test_gpu_parity.py -m gpu -x -q -k "beam or greedy" > gpurun_out/beam_tests.log 2>&1
tail -15 gpurun_out/beam_tests.log
