#!/bin/bash
for v in "PGEOF_KNN_TWO_LEVEL=0 PGEOF_KNN_LOCK=1" "PGEOF_KNN_TWO_LEVEL=0 PGEOF_KNN_LOCK=0" "PGEOF_KNN_TWO_LEVEL=0 PGEOF_KNN_FLAGS=2" "PGEOF_KNN_TWO_LEVEL=1"; do
  env $v python tools/lidar_probe.py 2>&1 | grep -E "lidar kNN" | tail -1 | tr '\n' ' '; echo " $v"
done
PGEOF_KNN_TWO_LEVEL=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"knn_tile|knn_slow|cell_count|scatter_kernel" python tools/lidar_probe.py 2>/dev/null | grep -E "knn_tile|knn_slow|cell_count|scatter" | tail -8 | awk -F"\",\"" '{print $5, $NF}' | cut -c1-120
