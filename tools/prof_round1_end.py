"""ncu workload, end of round 1: one launch set of the kernels added / reworked in the second half of the round:
batched MODE-DOTA cache step (cfg 5), DOTA.update inverse (D=512), kNN grouping with histogram selection (cfg 5),
cluster FPS (one 10 000-point cloud), DOTA fit (cfg 1), attention operand preparation."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200.engine import MultiStreamModeDota
from uniadapter_b200.gemm import attention_tf32x3
from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
dev = torch.device("cuda:0")
CFG = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
g = torch.Generator().manual_seed(1)
# cache step, cfg 5
K, M, D, B = 55, 8, 1024, 64
text = synthetic_text_features(K, D, 0).to(dev)
m = MultiStreamModeDota(CFG, D, K, text, M, 1, dev)
x = torch.nn.functional.normalize(torch.randn(1, B, D, device=dev), dim=-1)
gam = torch.softmax(100 * x @ text.t(), -1).contiguous()
xp = x.mean(1, keepdim=True).contiguous()
# DOTA, cfg 1
dota = ua.DOTA(CFG, 512, 40, torch.full((512, 40), 0.001), device=dev)
x1 = torch.nn.functional.normalize(torch.randn(1, 512, device=dev), dim=-1)
y1 = torch.softmax(torch.randn(1, 40, device=dev), 1)
# tokenizer
xyz = unit_sphere_clouds(64, 1024, g).to(dev)
rgb = torch.rand(64, 1024, 3, generator=g).to(dev)
_, centers = ua.fps_sample(xyz, 512, None)
big = unit_sphere_clouds(1, 10000, g).to(dev)
qkv = torch.randn(30 * 513, 3 * 384, generator=g).to(dev)
for _ in range(2):
    m.step(xp, x, gam)
    dota.fit(x1, y1); dota.update()
    ua.knn_group(xyz, centers, 64, rgb)
    ua.fps_sample(big, 512, None)
    attention_tf32x3(qkv, 30, 513, 6)
torch.cuda.synchronize()
print("ok")
