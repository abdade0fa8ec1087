#!/bin/bash
# round 2, visit x: ncu --set full of the shipped single-pair fill kernel (45000x45000), plain run first
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/bench_configs.py --configs square 2>&1 | cut -c1-170 || exit 1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fill_kernel --launch-count 2 -f -o $OUT/r02x_fill python tools/bench_configs.py --configs square > $OUT/r02x_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/r02x_ncu.log
