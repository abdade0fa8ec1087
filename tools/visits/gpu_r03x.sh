#!/bin/bash
# visit 3x: ncu --set full of the final batch fill kernel (band-major, one launch), plain run first
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/bench_configs.py --configs batch 2>&1 | cut -c1-170 || exit 1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fill_kernel --launch-count 1 -f -o $OUT/r03x_batch_fill python tools/bench_configs.py --configs batch > $OUT/r03x_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/r03x_ncu.log
