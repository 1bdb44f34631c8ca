#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2aa}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
CONFIGS=3 python tools/configs_probe.py 2>&1 | tail -6
PGEOF_RADIUS_TILE=0 CONFIGS=3 python tools/configs_probe.py 2>&1 | head -2
PGEOF_RADIUS_TILE=0 PGEOF_GENERIC_XF=1 CONFIGS=3 python tools/configs_probe.py 2>&1 | head -2
timeout 300 python tools/fuzz_search.py 150 2>&1 | tail -1
