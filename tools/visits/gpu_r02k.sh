#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q 2>&1 | tail -4
timeout 600 python tools/bench_configs.py --configs score,score_batch,batch,square 2>&1 | tee $OUT/r02k_configs.log
