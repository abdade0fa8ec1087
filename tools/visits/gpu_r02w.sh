#!/bin/bash
# round 2, visit w: single pairs on the batch geometry (64-row strips, many CTAs per SM), with and without half skew
set -u
for v in ss sshs; do for cfg in "32 1" "32 2" "64 2"; do set -- $cfg; echo "== $v KT=$1 WPC=$2";
  SWB_LIB=build/libswb200_$v.so SWB_KT=$1 SWB_WPC=$2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "random_shapes or medium" 2>&1 | tail -1
  SWB_LIB=build/libswb200_$v.so SWB_KT=$1 SWB_WPC=$2 timeout 600 python tools/bench_configs.py --configs square,big,batch 2>&1 | cut -c1-150; done; done
