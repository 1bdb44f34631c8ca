#!/bin/bash
# parameter sweep on the default bench: one line per setting
OUT=gpurun_out; TAG=${1:-r2ao}; shift
for v in "$@"; do
  env $v timeout 300 python bench.py ${BENCH_ARGS:---steps 6} --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v', round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"
done
