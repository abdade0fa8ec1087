#!/bin/bash
set -u
for v in pm bm; do for w in 1 2; do echo "== $v wpc=$w"; SWB_LIB=build/libswb200_$v.so SWB_WPC=$w timeout 300 python tools/bench_configs.py --configs batch 2>&1 | cut -c1-150; done; done
