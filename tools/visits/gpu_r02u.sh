#!/bin/bash
# round 2, visit u: poll granularity of the strip above (steps per poll): 8 (shipped) / 4 / 2
set -u
for v in w4 w2 w2all; do echo "== $v"; SWB_LIB=build/libswb200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "random_shapes or score_only or medium" 2>&1 | tail -1; SWB_LIB=build/libswb200_$v.so timeout 600 python tools/bench_configs.py --configs square,score,big,skew,skewT,batch,score_batch 2>&1 | cut -c1-150; done
