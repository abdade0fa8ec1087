#!/bin/bash
# visit 3r (2 GPUs): bench --gpus 2 with default flags after the buffer-release fix; strip tests
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r03r_bench2.json 2> $OUT/r03r_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03r_bench2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","serial","fill_only","kernel_ms_per_rank","parity")}); print(d['e2e']); print(d['secondary'])
PY
grep -v "OMP_NUM\|\*\*\*" $OUT/r03r_bench2.err | tail -4
timeout 900 python -m pytest tests/test_gpu_strips.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
