#!/bin/bash
# round 2, visit p: wide family + pitch / wave sweep
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02p_pytest.log
timeout 900 python tools/bench_configs.py --configs batch,score_batch,skew,skewT,pitch 2>&1 | tee $OUT/r02p_configs.log
