#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2n}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "knn or upstream or metric or switch or clip or local or csr or multiscale or optimal" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
for v in "PGEOF_KNN_TILE128=1" "PGEOF_KNN_TILE128=0"; do
  env $v PGEOF_KNN_STATS=1 timeout 300 python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 >/dev/null | grep "pgeof knn tile" | tail -1 | cut -c1-330
  env $v timeout 300 python bench.py --config C4 --steps 5 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v C4', 'step %.2f'%d['ms_per_step'], d['roofline']['all_kernels'])"
done
timeout 300 python tools/fuzz_search.py 150 2>&1 | tail -2
