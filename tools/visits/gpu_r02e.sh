#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in r3k7 r3k8 r2k7; do
  echo "== lib $v"
  export SWB_LIB=build/libswb200_$v.so
  timeout 600 python tools/bench_configs.py --configs square,big 2>&1 | tee -a $OUT/r02e_configs.log
done
export SWB_LIB=build/libswb200_r3k7.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "random_shapes or builtin or small_dumps or medium" 2>&1 | tail -3
echo "== grouptrace r3k7"; SHAPE=45000 SWB_LIB=build/libswb200_r3k7gt.so timeout 300 python tools/grouptrace.py 2 96 2>&1 | tee $OUT/r02e_grouptrace_r3k7.log
