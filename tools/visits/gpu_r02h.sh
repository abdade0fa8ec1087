#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02h_bench2.json 2> $OUT/r02h_bench2.err; echo "bench2 rc=$?"; cat $OUT/r02h_bench2.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/r02h_bench2.err | tail -8
