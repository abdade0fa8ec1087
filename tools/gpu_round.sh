#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list + full capture of the fill kernel.
# Usage (under gpurun): bash tools/gpu_round.sh [tag] [skip_tests]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.log 2>&1
free -g | head -2 >> $OUT/gpu_$TAG.log; nproc >> $OUT/gpu_$TAG.log
if [ "${2:-}" != "skip_tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
fi
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1; echo "bench ref rc=$?"
cat $OUT/bench_ref_$TAG.json
BCMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $BCMD > $OUT/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv $BCMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 600 $BCMD > $OUT/plain2_$TAG.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 1 -c 1 \
    -f -o $OUT/prof_fill_$TAG $BCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
