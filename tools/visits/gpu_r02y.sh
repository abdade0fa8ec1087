#!/bin/bash
# round 2, visit y (2 GPUs): packed transfer with row base: tests, N=1 e2e, N=2 bench with packed e2e
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_pack.py tests/test_gpu_parity.py tests/test_gpu_strips.py tests/test_gpu_multi.py -m gpu -x -q > $OUT/r02y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02y_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-secondary --steps 3 > $OUT/r02y_bench1.json 2> $OUT/r02y_bench1.err; echo "bench1 rc=$?"; python -c "import json; d=json.loads(open('$OUT/r02y_bench1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-secondary > $OUT/r02y_bench2.json 2> $OUT/r02y_bench2.err; echo "bench2 rc=$?"; python -c "import json; d=json.loads(open('$OUT/r02y_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['parity'], d['e2e'])"; grep -v "OMP_NUM\|\*\*\*" $OUT/r02y_bench2.err | tail -5
