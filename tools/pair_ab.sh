#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2w}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "knn or upstream or metric or radius or fused or switch or clip or local or csr" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
for v in "PGEOF_KNN_ROLLED=0" "PGEOF_KNN_ROLLED=1"; do
  env $v timeout 300 python bench.py --steps 8 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v', 'step %.2f'%d['ms_per_step'], 'knn %.2f'%d['roofline']['all_kernels']['knn_search']['ms'])"
  env $v timeout 300 ncu --metrics gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,smsp__issue_active.avg.per_cycle_active,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio --clock-control none -k regex:knn_tile -s 3 -c 1 --csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep -E "knn_tile" | awk -F'","' '{print "    ", $(NF-2), $(NF)}'
done
python tools/lidar_probe.py 2>&1 | tail -1
timeout 300 python bench.py --config C4 --steps 5 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('C4', 'step %.2f'%d['ms_per_step'], d['roofline']['all_kernels'])"
timeout 200 python tools/fuzz_search.py 100 2>&1 | tail -1
