#!/bin/bash
TAG=${1:-r2as}; OUT=gpurun_out; mkdir -p $OUT
for v in "PGEOF_FEATURES_TEX=2" "PGEOF_FEATURES_TEX=1"; do
  for c in "" "--config C4"; do
  env $v timeout 600 python bench.py $c --steps 4 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v $c', round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"
  done
done
