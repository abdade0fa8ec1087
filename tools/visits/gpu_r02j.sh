#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $OUT/r02j_gpu.log; free -g | head -2 >> $OUT/r02j_gpu.log; nproc >> $OUT/r02j_gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/r02j_bench8.json 2> $OUT/r02j_bench8.err; echo "bench8 rc=$?"; cat $OUT/r02j_bench8.json | cut -c1-400; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/r02j_bench8.err | tail -8
