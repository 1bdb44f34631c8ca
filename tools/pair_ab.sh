#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2i}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "optimal or switch" > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log
for v in "PGEOF_OPTIMAL_SCAN=1"; do
  env $v timeout 600 python bench.py --config C5 --points 10000000 --steps 3 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v C5@10M', 'step %.2f'%d['ms_per_step'], d['roofline']['all_kernels'])"
done
