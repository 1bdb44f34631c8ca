#!/bin/bash
# multiscale A/B: tests, then C4 with the two-pass (default) and the one-pass kernel
TAG=${1:-r2ap}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "multiscale or switch or c4" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
for v in "PGEOF_MULTISCALE_SPLIT=1" "PGEOF_MULTISCALE_SPLIT=0"; do
  env $v timeout 600 python bench.py --config C4 --steps 4 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v', d['ms_per_step'], d['roofline']['all_kernels'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"multiscale" -c 6 python bench.py --config C4 --steps 1 --warmup 1 --no-e2e --no-cpu 2>/dev/null | grep multiscale | awk -F'","' '{print substr($5,1,50), $NF}'
