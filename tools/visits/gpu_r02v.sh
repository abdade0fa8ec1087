#!/bin/bash
# round 2, visit v: final build: full GPU suite, all configs, traces, bench (own arm + reference arm), then ncu
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02v_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python tools/bench_configs.py --configs square,score,batch,score_batch,big,skew,skewT 2>&1 | tee $OUT/r02v_configs.log
echo "== trace"; SWB_LIB=build/libswb200_trace.so timeout 300 python tools/trace.py --shape 45000x45000 2>&1 | tail -14 | tee $OUT/r02v_trace.log
echo "== grouptrace"; SWB_LIB=build/libswb200_gt.so timeout 300 python tools/grouptrace.py 2>&1 | tail -6 | tee $OUT/r02v_grouptrace.log
echo "== bench"; timeout 900 python bench.py > $OUT/r02v_bench.json 2> $OUT/r02v_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/r02v_bench.json; tail -3 $OUT/r02v_bench.err
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02v_bench_ref.json 2> $OUT/r02v_bench_ref.err; echo "ref rc=$?"; cut -c1-600 $OUT/r02v_bench_ref.json
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02v_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $OUT/r02v_ncu_launches.log 2>&1; echo "rc=$?"
