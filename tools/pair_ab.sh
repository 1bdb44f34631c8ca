#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2s}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "chunked or features or multiscale or optimal or golden or torch or abi or ctypes" > $OUT/${TAG}_pytest.log 2>&1; tail -5 $OUT/${TAG}_pytest.log
for v in "PGEOF_HOST_CHUNK_MB=0" "PGEOF_HOST_CHUNK_MB=128" "PGEOF_HOST_CHUNK_MB=256" "PGEOF_HOST_CHUNK_MB=64"; do
  env $v PGEOF_HOST_TRACE=1 timeout 300 python bench.py --steps 3 --e2e-steps 4 --no-cpu > $OUT/${TAG}_ab.json 2> $OUT/${TAG}_ab.err
  grep "pgeof host" $OUT/${TAG}_ab.err | tail -2
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v', 'e2e %.2f ms'%d['e2e']['ms_per_step'], d['e2e']['step_ms_min_median_max'])"
done
