cat > /tmp/rad.py <<PY
import sys; sys.path.insert(0, "/root/repo")
import torch, pgeof
from point_geometric_features_b200 import synth
t = torch.from_numpy(synth.lidar_like_cloud(10_000_000, seed=0)).cuda()
for _ in range(2):
    idx, d2 = pgeof.radius_search(t, t, 0.2, 64)
torch.cuda.synchronize()
PY
PGEOF_KNN_STATS=1 python /tmp/rad.py 2>&1 | grep "pgeof knn tile" | head -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rad_launches.csv python /tmp/rad.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/rad_launches.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[len(rows)//2:]: print(r[0], r[4][:70], r[-1])
PY
