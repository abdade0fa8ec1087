#!/bin/bash
# visit 3m: band-major batches: full GPU suite + configs
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r03m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r03m_pytest.log
timeout 900 python tools/bench_configs.py --configs square,score,batch,score_batch,big,skew,skewT 2>&1 | tee $OUT/r03m_configs.log | cut -c1-170
