#!/bin/bash
# visit 3i: row-wise backtrack walker: parity, time
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_strips.py tests/test_gpu_cli.py tests/test_pack.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bt_time.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -2
