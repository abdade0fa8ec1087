#!/bin/bash
# visit 3g (2 GPUs): overlapped strip fills at N=2, with and without overlap
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 --no-secondary --no-e2e "$@" > $OUT/r03g_$tag.json 2> $OUT/r03g_$tag.err; echo "$tag rc=$?"; python -c "import json; d=json.loads(open('$OUT/r03g_$tag.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['serial']['ms_per_step'], d['fill_only']['ms_max_over_ranks'], d['parity'])"; grep -v "OMP_NUM\|\*\*\*" $OUT/r03g_$tag.err | tail -3; }
run overlap
run nooverlap --no-overlap
