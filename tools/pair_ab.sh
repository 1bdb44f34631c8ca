#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2j}
for v in 4 5 6 8; do
  PGEOF_OPTIMAL_SCAN=$v timeout 600 python bench.py --config C5 --points 10000000 --steps 3 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v C5@10M', 'step %.2f'%d['ms_per_step'], d['roofline']['all_kernels']['optimal'])"
done
