#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
unset SWB_LIB
timeout 300 python tools/bench_configs.py --configs square,big,score 2>&1 | tee -a $OUT/r02m_configs.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "== grouptrace"; SHAPE=45000 SWB_LIB=build/libswb200_gt.so timeout 300 python tools/grouptrace.py 2 96 2>&1 | tail -6
