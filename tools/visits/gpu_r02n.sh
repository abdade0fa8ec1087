#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
BCMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary"
timeout 600 $BCMD > $OUT/r02n_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 2 -c 1 -f -o $OUT/r02n_prof_fill $BCMD > $OUT/r02n_ncu_full.log 2>&1
echo "ncu full rc=$?"
echo "== trace"; SWB_LIB=build/libswb200_trace.so timeout 300 python tools/trace.py --shape 45000x45000 2>&1 | tail -12 | tee $OUT/r02n_trace.log
echo "== grouptrace"; SHAPE=45000 SWB_LIB=build/libswb200_gt.so timeout 300 python tools/grouptrace.py 2 96 2>&1 | tail -9 | tee $OUT/r02n_grouptrace.log
