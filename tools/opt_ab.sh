#!/bin/bash
# optimal-k kernel: CTAs per SM sweep for the closed-form filter + one ncu capture
TAG=${1:-r2ai}; OUT=gpurun_out; mkdir -p $OUT
for v in 4 5 6 8; do
  PGEOF_OPTIMAL_CTAS=$v timeout 600 python bench.py --config C5 --points 10000000 --steps 3 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('ctas=$v', d['ms_per_step'], d['roofline']['all_kernels']['optimal'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"optimal_scan" -s 1 -c 1 -o $OUT/${TAG}_opt -f python bench.py --config C5 --points 2000000 --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_opt.log 2>&1; tail -1 $OUT/${TAG}_ncu_opt.log
