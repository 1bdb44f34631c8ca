#!/bin/bash
for sp in 1.2 1.7 2.5; do for co in 2 3 4 6; do
  PGEOF_KNN_SPLIT=$sp PGEOF_KNN_COARSE=$co PGEOF_KNN_STATS=1 python tools/lidar_probe.py 2>&1 | grep -E "deferred|lidar kNN" | tail -2 | tr '\n' ' '; echo " split=$sp coarse=$co"
done; done
