#!/bin/bash
# round 2, visit f: dual-geometry build: full GPU test suite, configs, bench
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/r02f_pytest.log
timeout 900 python tools/bench_configs.py --configs square,score,batch,score_batch,big,skew,skewT 2>&1 | tee $OUT/r02f_configs.log
echo "== trace"; SWB_LIB=build/libswb200_trace.so timeout 300 python tools/trace.py --shape 45000x45000 2>&1 | tail -14 | tee $OUT/r02f_trace.log
echo "== bench"; timeout 900 python bench.py > $OUT/r02f_bench.json 2> $OUT/r02f_bench.err; echo "bench rc=$?"; cat $OUT/r02f_bench.json; tail -5 $OUT/r02f_bench.err
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02f_bench_ref.json 2> $OUT/r02f_bench_ref.err; echo "ref rc=$?"; cat $OUT/r02f_bench_ref.json
