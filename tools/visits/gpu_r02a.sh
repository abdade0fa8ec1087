#!/bin/bash
# round 2, visit a: parity of the profile kernel + kernel times of variants + in-situ traces
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r02a_gpu.log 2>&1
nproc >> $OUT/r02a_gpu.log; free -g | head -2 >> $OUT/r02a_gpu.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_strips.py -x -q > $OUT/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/r02a_pytest.log
for v in "" build/libswb200_r1.so build/libswb200_w4.so; do
  echo "== lib ${v:-product}"
  if [ -n "$v" ]; then export SWB_LIB=$v; else unset SWB_LIB; fi
  timeout 600 python tools/bench_configs.py --configs square,score 2>&1 | tee -a $OUT/r02a_configs.log
done
unset SWB_LIB
timeout 900 python tools/bench_configs.py --configs big,batch,score_batch,skew,skewT 2>&1 | tee -a $OUT/r02a_configs.log
echo "== grouptrace"; SHAPE=45000 SWB_LIB=build/libswb200_gt.so timeout 300 python tools/grouptrace.py 2 2>&1 | tee $OUT/r02a_grouptrace.log
echo "== trace"; SWB_LIB=build/libswb200_trace.so timeout 300 python tools/trace.py --shape 45000x45000 2>&1 | tee $OUT/r02a_trace.log
echo "== large"; timeout 1200 python -m pytest tests/test_gpu_large.py -x -q > $OUT/r02a_pytest_large.log 2>&1; echo "pytest large rc=$?"; tail -15 $OUT/r02a_pytest_large.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02a_bench.json 2> $OUT/r02a_bench.err; echo "bench rc=$?"; cat $OUT/r02a_bench.json; tail -5 $OUT/r02a_bench.err
