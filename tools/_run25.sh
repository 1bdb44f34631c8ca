cat > /tmp/ms.py <<PY
import sys; sys.path.insert(0, "/root/repo")
import torch, pgeof
from point_geometric_features_b200 import synth
t = torch.from_numpy(synth.uniform_cloud(10_000_000, seed=0)).cuda()
idx, _ = pgeof.knn_search(t, t, 100)
ptr = (torch.arange(t.shape[0] + 1, device="cuda", dtype=torch.int64) * 100).to(torch.uint32)
for _ in range(2):
    f = pgeof.compute_features_multiscale(t, idx.view(-1), ptr, [10, 20, 50, 100])
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"multiscale_direct" -s 1 -c 1 -o gpurun_out/ms_prof python /tmp/ms.py > /dev/null 2>&1
