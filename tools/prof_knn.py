"""ncu workload: kNN grouping at the cfg 5 shape (B=64 clouds of 1024 points, 512 centres, k=64, colour concat)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200.streams import unit_sphere_clouds
dev = torch.device("cuda:0")
B, N, G, k = int(os.environ.get("PB", 64)), 1024, 512, int(os.environ.get("PKNN", 64))
g = torch.Generator().manual_seed(1)
xyz = unit_sphere_clouds(B, N, g).to(dev)
rgb = torch.rand(B, N, 3, generator=g).to(dev)
_, centers = ua.fps_sample(xyz, G, None)
for _ in range(3):
    ua.knn_group(xyz, centers, k, rgb)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): ua.knn_group(xyz, centers, k, rgb)
e.record(); torch.cuda.synchronize()
print("knn_group", round(s.elapsed_time(e) / 20 * 1e3, 1), "us per call (back to back)")
