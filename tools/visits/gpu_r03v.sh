#!/bin/bash
# visit 3v: writer knobs on the batch configuration (band-major build)
set -u
for v in ws32 ws256 wd2 wd8 np; do echo "== $v"; SWB_LIB=build/libswb200_$v.so timeout 300 python tools/bench_configs.py --configs batch,square 2>&1 | cut -c1-150; done
