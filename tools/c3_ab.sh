#!/bin/bash
TAG=${1:-r2an}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "radius or csr or c3 or adaptor or beyond_512" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --config C3 --steps 5 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('C3', d['ms_per_step'], d['roofline']['all_kernels'])"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches_C3.csv python bench.py --config C3 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; tail -22 $OUT/${TAG}_launches_C3.csv | grep -E "knn_tile|search_list|padded" | awk -F'","' '{print substr($5,1,60), $NF}'
