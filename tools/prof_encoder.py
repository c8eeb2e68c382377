"""ncu workload: eager encoder forwards of the bench shape (15 clouds x 1024 points, ULIP-2 PointBERT)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uniadapter_b200.encoders import build_encoder
from uniadapter_b200.streams import unit_sphere_clouds
dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 15
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
enc = build_encoder('ulip', 0, dev)
pc = unit_sphere_clouds(S, 1024, torch.Generator().manual_seed(1)).to(dev)
with torch.no_grad():
    for _ in range(n):
        enc(pc)
torch.cuda.synchronize()
print("ok")
