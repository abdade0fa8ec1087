#!/bin/bash
# visit 3a: large-pair overlap in swb_fill_pairs_async + overlapped bench steps
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_bench_contract.py -m gpu -x -q > $OUT/r03a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r03a_pytest.log
echo "== bench"; timeout 900 python bench.py > $OUT/r03a_bench.json 2> $OUT/r03a_bench.err; echo "bench rc=$?"; tail -3 $OUT/r03a_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03a_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","serial","parity","gpu_launches")})
print(d['roofline']); print(d['e2e']); print(d['secondary'].get('large_pairs')); print(d['config']['pipeline'])
PY
