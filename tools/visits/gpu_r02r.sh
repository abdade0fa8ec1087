#!/bin/bash
# round 2, visit r: column-strip bench on 8, 4, 2 GPUs of one box with the half-skew build
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 5 --warmup 3 "$@" > $OUT/r02r_bench$n.json 2> $OUT/r02r_bench$n.err; echo "bench$n rc=$?"; python - $OUT/r02r_bench$n.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","serial","fill_only","kernel_ms_per_rank","parity")}, d.get("e2e"))
PY
}
run 8
run 4 --no-e2e --no-secondary
run 2 --no-e2e --no-secondary
