#!/bin/bash
# visit 4b: final state: full GPU suite, smoke, bench with default flags
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r04b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r04b_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > $OUT/r04b_bench.json 2> $OUT/r04b_bench.err; echo "bench rc=$?"; python -c "import json; d=json.loads(open('$OUT/r04b_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['serial']['value'], d['roofline']['frac'], d['roofline']['sustained']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['parity'], d['e2e']['parity'])"
