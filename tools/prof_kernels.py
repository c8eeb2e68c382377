"""Small fixed workload for ncu captures: a few launches of each hot kernel at BASELINE shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
if which in ("all", "modedota"):
    K, M, D = 1156, 8, 1024
    text = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
    model = ua.DOTA_mix(cfg, D, K, text.t().contiguous(), num_modes=M, device=dev)
    x = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
    g = torch.softmax(100 * x @ text.t(), 1)
    for _ in range(4):
        model.predict_then_fit(x, x, g)
if which in ("all", "tok"):
    for (B, N, G, k) in [(64, 1024, 512, 32), (1, 10000, 512, 64)]:
        xyz = torch.from_numpy(synth.cloud(B, N, 1)).to(dev)
        for _ in range(2):
            _, cen = ua.fps_sample(xyz, G, None)
            ua.knn_group(xyz, cen, k)
if which in ("all", "dota"):
    K, D = 40, 512
    model = ua.DOTA(cfg, D, K, torch.full((D, K), 0.001), device=dev)
    x = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
    y = torch.softmax(torch.randn(1, K, device=dev), 1)
    for _ in range(3):
        model.fit(x, y)
        model.predict(x.half())
torch.cuda.synchronize()
print("ok")
