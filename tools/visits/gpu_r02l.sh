#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in pf2 ""; do
  echo "== lib ${v:-product}"
  if [ -n "$v" ]; then export SWB_LIB=build/libswb200_$v.so; else unset SWB_LIB; fi
  timeout 300 python tools/bench_configs.py --configs square,big,score 2>&1 | tee -a $OUT/r02l_configs.log
done
export SWB_LIB=build/libswb200_pf2.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "== grouptrace pf2"; SHAPE=45000 SWB_LIB=build/libswb200_pf2gt.so timeout 300 python tools/grouptrace.py 2 96 2>&1 | tail -8
