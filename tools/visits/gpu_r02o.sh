#!/bin/bash
# round 2, visit o: half-skew build: full GPU suite, configs, bench, reference arm
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/r02o_pytest.log
timeout 900 python tools/bench_configs.py --configs square,score,batch,score_batch,big,skew,skewT 2>&1 | tee $OUT/r02o_configs.log
echo "== bench"; timeout 900 python bench.py > $OUT/r02o_bench.json 2> $OUT/r02o_bench.err; echo "bench rc=$?"; cut -c1-600 $OUT/r02o_bench.json; tail -5 $OUT/r02o_bench.err
