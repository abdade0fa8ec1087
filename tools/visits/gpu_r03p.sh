#!/bin/bash
# round 2, visit 3p: final build: full GPU suite, all configs, traces, bench (own arm + reference arm), then ncu
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r03p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r03p_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python tools/bench_configs.py --configs square,score,batch,score_batch,big,skew,skewT 2>&1 | tee $OUT/r03p_configs.log
echo "== bench"; timeout 900 python bench.py > $OUT/r03p_bench.json 2> $OUT/r03p_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/r03p_bench.json; tail -3 $OUT/r03p_bench.err
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r03p_bench_ref.json 2> $OUT/r03p_bench_ref.err; echo "ref rc=$?"; cut -c1-600 $OUT/r03p_bench_ref.json
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r03p_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $OUT/r03p_ncu_launches.log 2>&1; echo "rc=$?"
