#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 build/ubench_dp4a 2>&1 | tee $OUT/r02c_ubench_dp4a.log
