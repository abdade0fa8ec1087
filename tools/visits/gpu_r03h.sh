#!/bin/bash
# visit 3h (8 GPUs): overlapped strip fills at N=8 and N=4
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { n=$1; tag=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) bench.py --gpus $n --steps 6 --warmup 3 --no-secondary --no-e2e "$@" > $OUT/r03h_$tag.json 2> $OUT/r03h_$tag.err; echo "$tag rc=$?"; python -c "import json; d=json.loads(open('$OUT/r03h_$tag.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['serial']['ms_per_step'], d['fill_only']['ms_max_over_ranks'], d['kernel_ms_per_rank'], d['parity'])"; grep -v "OMP_NUM\|\*\*\*" $OUT/r03h_$tag.err | tail -3; }
run 8 n8_overlap
run 8 n8_nooverlap --no-overlap
run 4 n4_overlap
