#!/bin/bash
# round 2, visit t: ncu --set full of the batch fill (65536 x 256x256, full H+P) -- plain run first
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/bench_configs.py --configs batch 2>&1 | cut -c1-170 || exit 1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fill_kernel --launch-count 2 -f -o $OUT/r02t_batch_fill python tools/bench_configs.py --configs batch > $OUT/r02t_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/r02t_ncu.log
ls -la $OUT/r02t_batch_fill.ncu-rep
