#!/bin/bash
# visit 3u: merged forms for the batch geometry only: full GPU suite + configs + bench
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r03u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r03u_pytest.log
timeout 900 python tools/bench_configs.py --configs square,score,batch,score_batch,big,skew,skewT 2>&1 | tee $OUT/r03u_configs.log | cut -c1-170
timeout 600 python tools/large_pairs_time.py 2>&1 | head -5
timeout 900 python bench.py > $OUT/r03u_bench.json 2> $OUT/r03u_bench.err; echo "bench rc=$?"; python -c "import json; d=json.loads(open('$OUT/r03u_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['serial']['value'], d['roofline']['frac'], d['roofline']['sustained']['frac'], d['e2e']['value'], d['gpu_launches'], d['parity'])"
