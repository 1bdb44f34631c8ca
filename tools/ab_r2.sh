#!/bin/bash
# A/B of the tile kernel's scan on one GPU: search tests + fuzz, then the default bench with and without a switch.
# Usage: tools/ab_r2.sh TAG "ENV=.. ENV=.." ["ENV=.."...]
TAG=${1:-r2x}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "knn or radius or search or fused or switch or metric" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
[ -n "$FUZZ" ] && timeout 300 python tools/fuzz_search.py 2>&1 | tail -2
for v in "PGEOF_NOP=1" "$@"; do
  env $v timeout 300 python bench.py --steps 10 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v', d['ms_per_step'], d['roofline']['all_kernels'])"
done
