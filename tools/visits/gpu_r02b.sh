#!/bin/bash
# round 2, visit b: profile-kernel variants (smem ring / L1 loads / compare), parity, traces
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r02b_gpu.log 2>&1
nproc >> $OUT/r02b_gpu.log; free -g | head -2 >> $OUT/r02b_gpu.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/r02b_pytest.log 2>&1; echo "pytest parity rc=$?"; tail -8 $OUT/r02b_pytest.log
for v in "" cmp v1 v1np w4; do
  echo "== lib ${v:-product}"
  if [ -n "$v" ]; then export SWB_LIB=build/libswb200_$v.so; else unset SWB_LIB; fi
  timeout 600 python tools/bench_configs.py --configs square,score,batch 2>&1 | tee -a $OUT/r02b_configs.log
done
unset SWB_LIB
timeout 900 python tools/bench_configs.py --configs big,score_batch,skew,skewT 2>&1 | tee -a $OUT/r02b_configs.log
echo "== grouptrace product"; SHAPE=45000 SWB_LIB=build/libswb200_gt.so timeout 300 python tools/grouptrace.py 2 2>&1 | tee $OUT/r02b_grouptrace.log
echo "== grouptrace compare"; SHAPE=45000 SWB_LIB=build/libswb200_gtcmp.so timeout 300 python tools/grouptrace.py 2 2>&1 | tee $OUT/r02b_grouptrace_cmp.log
echo "== trace"; SWB_LIB=build/libswb200_trace.so timeout 300 python tools/trace.py --shape 45000x45000 2>&1 | tee $OUT/r02b_trace.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_strips.py tests/test_gpu_cli.py -x -q > $OUT/r02b_pytest2.log 2>&1; echo "pytest multi rc=$?"; tail -15 $OUT/r02b_pytest2.log
echo "== large"; timeout 1200 python -m pytest tests/test_gpu_large.py -x -q > $OUT/r02b_pytest_large.log 2>&1; echo "pytest large rc=$?"; tail -15 $OUT/r02b_pytest_large.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02b_bench.json 2> $OUT/r02b_bench.err; echo "bench rc=$?"; cat $OUT/r02b_bench.json; tail -5 $OUT/r02b_bench.err
echo "== bt2 (jump-table backtrack) parity + timing"
SWB_LIB=build/libswb200_bt2.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q > $OUT/r02b_pytest_bt2.log 2>&1; echo "pytest bt2 rc=$?"; tail -8 $OUT/r02b_pytest_bt2.log
for v in "" bt2; do
  if [ -n "$v" ]; then export SWB_LIB=build/libswb200_$v.so; else unset SWB_LIB; fi
  timeout 300 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary --no-pipeline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('lib ${v:-product}: step', d['ms_per_step'], 'fill', d['fill_ms'], 'backtrack est', d['backtrack_ms_est'], d['result'], d['parity'])
"
done
unset SWB_LIB
