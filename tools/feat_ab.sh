#!/bin/bash
# A/B of the feature-kernel layouts and L2 policies (one gpurun call): bench lines + dram counters of one launch each
OUT=gpurun_out; TAG=${1:-r2b}
mkdir -p $OUT
for v in "PGEOF_FEATURES_RANK=0 PGEOF_FEATURES_POLICY=0" "PGEOF_FEATURES_RANK=0 PGEOF_FEATURES_POLICY=1" "PGEOF_FEATURES_RANK=0 PGEOF_FEATURES_POLICY=5" \
         "PGEOF_FEATURES_RANK=1 PGEOF_FEATURES_POLICY=0" "PGEOF_FEATURES_RANK=1 PGEOF_FEATURES_POLICY=1" "PGEOF_FEATURES_RANK=1 PGEOF_FEATURES_POLICY=3" \
         "PGEOF_FEATURES_RANK=1 PGEOF_FEATURES_POLICY=7" "PGEOF_FEATURES_RANK=1 PGEOF_FEATURES_POLICY=2"; do
  env $v timeout 300 python bench.py --steps 8 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null
  python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); k=d['roofline']['all_kernels']; print('$v', 'step %.2f'%d['ms_per_step'], 'feat %.3f'%k['features']['ms'], 'prepass %.3f'%k['row_order']['ms'])"
  env $v timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:features_direct -s 3 -c 1 --csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep -E "features_direct" | awk -F'","' '{print "    ", $(NF-2), $(NF)}' 
done
