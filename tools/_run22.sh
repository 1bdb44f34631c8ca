cat > /tmp/rad.py <<PY
import sys; sys.path.insert(0, "/root/repo")
import torch, pgeof
from point_geometric_features_b200 import synth
t = torch.from_numpy(synth.lidar_like_cloud(10_000_000, seed=0)).cuda()
for _ in range(2):
    idx, d2 = pgeof.radius_search(t, t, 0.2, 64)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel" -s 1 -c 1 -o gpurun_out/rad_prof python /tmp/rad.py > /dev/null 2>&1
