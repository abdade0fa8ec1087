#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/bench_configs.py --configs square,batch,score_batch 2>&1 | cut -c1-170
timeout 600 python tools/large_pairs_time.py 2>&1 | head -5
