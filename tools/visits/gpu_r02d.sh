#!/bin/bash
# round 2, visit d: selector/dp4a fill kernel: parity, variants, traces
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/r02d_pytest.log 2>&1; echo "pytest parity rc=$?"; tail -8 $OUT/r02d_pytest.log
for v in "" late w4 w4late cmp; do
  echo "== lib ${v:-product}"
  if [ -n "$v" ]; then export SWB_LIB=build/libswb200_$v.so; else unset SWB_LIB; fi
  timeout 600 python tools/bench_configs.py --configs square,score,batch 2>&1 | tee -a $OUT/r02d_configs.log
done
unset SWB_LIB
timeout 900 python tools/bench_configs.py --configs big,score_batch,skew,skewT 2>&1 | tee -a $OUT/r02d_configs.log
echo "== grouptrace product"; SHAPE=45000 SWB_LIB=build/libswb200_gt.so timeout 300 python tools/grouptrace.py 2 2>&1 | tee $OUT/r02d_grouptrace.log
echo "== grouptrace late"; SHAPE=45000 SWB_LIB=build/libswb200_gtlate.so timeout 300 python tools/grouptrace.py 2 2>&1 | tee $OUT/r02d_grouptrace_late.log
echo "== trace"; SWB_LIB=build/libswb200_trace.so timeout 300 python tools/trace.py --shape 45000x45000 2>&1 | tail -14 | tee $OUT/r02d_trace.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_strips.py tests/test_gpu_cli.py tests/test_gpu_large.py -x -q > $OUT/r02d_pytest2.log 2>&1; echo "pytest multi+large rc=$?"; tail -8 $OUT/r02d_pytest2.log
