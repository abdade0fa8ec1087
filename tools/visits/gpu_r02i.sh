#!/bin/bash
# round 2, visit i: band-major batches; ncu launch list + full capture of the fill kernel
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q -k "batch or score" 2>&1 | tail -4
timeout 600 python tools/bench_configs.py --configs batch,score_batch,square 2>&1 | tee $OUT/r02i_configs.log
BCMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary"
timeout 600 $BCMD > $OUT/r02i_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches.csv $BCMD > $OUT/r02i_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 $BCMD > $OUT/r02i_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 2 -c 1 -f -o $OUT/r02_prof_fill $BCMD > $OUT/r02i_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -8
