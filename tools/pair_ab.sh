#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2y}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "knn or upstream or metric or radius or fused or switch or clip or local or csr or selected or c3 or c4 or c5 or randomised" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --config C4 --steps 5 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('C4', 'step %.2f'%d['ms_per_step'], d['roofline']['all_kernels'])"
python tools/lidar_probe.py 2>&1 | tail -1
timeout 300 python bench.py --config C3 --steps 5 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('C3', 'step %.2f'%d['ms_per_step'], d['roofline']['all_kernels'])"
timeout 300 python bench.py --steps 8 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('M', 'step %.2f'%d['ms_per_step'], d['roofline']['all_kernels'])"
timeout 300 python tools/fuzz_search.py 200 2>&1 | tail -1
