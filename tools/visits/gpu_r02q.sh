#!/bin/bash
# round 2, visit q: packed copy-back (tests + e2e in bench), also with SWB_PACKED_D2H=0 for the before/after
set -u
OUT=gpurun_out; mkdir -p $OUT
nproc; lscpu | grep -i "model name\|^CPU(s)\|numa" | head -6
timeout 900 python -m pytest tests/test_pack.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02q_pytest.log
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > $OUT/r02q_bench.json 2> $OUT/r02q_bench.err; echo "bench rc=$?"; tail -5 $OUT/r02q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['serial'], d['e2e'])
PY
for t in 4 8 16 32; do echo "== threads $t"; SWB_HOST_THREADS=$t timeout 600 python bench.py --no-cpu-baseline --no-secondary --steps 3 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['ms_per_step'], d['e2e']['value'])"; done
