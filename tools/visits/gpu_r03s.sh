#!/bin/bash
# visit 3s (8 GPUs): final bench --gpus 8 and --gpus 4 with default flags
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { n=$1; timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29570 + n)) bench.py --gpus $n --steps 5 --warmup 3 > $OUT/r03s_bench$n.json 2> $OUT/r03s_bench$n.err; echo "bench$n rc=$?"; python - $OUT/r03s_bench$n.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","serial","fill_only","kernel_ms_per_rank","parity")}); print(d.get("e2e")); print(d.get("secondary"))
PY
grep -v "OMP_NUM\|\*\*\*" $OUT/r03s_bench$n.err | tail -3; }
run 8
run 4
