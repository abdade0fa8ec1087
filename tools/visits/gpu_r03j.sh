#!/bin/bash
set -u
timeout 300 python tools/bt_time.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strips.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
