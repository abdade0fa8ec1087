#!/bin/bash
# visit 3q (2 GPUs): final build: whole GPU suite on a 2-GPU box (strip / NCCL / IPC tests run), bench --gpus 2 with default flags
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r03q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r03q_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r03q_bench2.json 2> $OUT/r03q_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03q_bench2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","serial","fill_only","kernel_ms_per_rank","parity")}); print(d['e2e']); print(d['secondary'])
PY
grep -v "OMP_NUM\|\*\*\*" $OUT/r03q_bench2.err | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-full 0 2>&1 | tail -2 | cut -c1-300
