#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2g}
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"optimal_scan" -s 1 -c 1 -o $OUT/${TAG}_opt -f python bench.py --config C5 --points 2000000 --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_opt.log 2>&1; tail -1 $OUT/${TAG}_ncu_opt.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel|features_direct" -s 6 -c 2 -o $OUT/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/${TAG}_ncu.log 2>&1; tail -1 $OUT/${TAG}_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
