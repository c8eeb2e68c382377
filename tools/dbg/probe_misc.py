import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import uniadapter_b200 as ua
from bench import L2Flush, median_us, CFG
from uniadapter_b200 import _lib
from uniadapter_b200.streams import unit_sphere_clouds
dev = torch.device("cuda:0")
flush = L2Flush(dev)
# ---- DOTA fit ------------------------------------------------------------------------------------------------
for K, D in [(40, 512), (40, 1024)]:
    a = ua.DOTA(CFG, D, K, torch.full((D, K), 0.001), device=dev)
    x = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
    y = torch.softmax(torch.randn(1, K, device=dev), 1)
    by = 8 * K * D * D + 4 * D * D
    for staged in (1, 0):
        _lib.set_tuning("dota_staged", staged)
        us = median_us(lambda: a.fit(x, y), flush)
        print(f"DOTA fit K={K} D={D} staged={staged}: {us:.1f} us  {by / us / 1e3:.0f} GB/s  frac {by / us / 1e3 / 6542.1:.3f}")
    _lib.set_tuning("dota_staged", 1)
# ---- FPS variants ---------------------------------------------------------------------------------------------
for B in (1, 64, 592):
    xyz = unit_sphere_clouds(B, 1024, torch.Generator().manual_seed(B)).to(dev)
    for pn2 in (False, True):
        us = median_us(lambda: ua.fps_sample(xyz, 512, None, pointnet2=pn2), flush, n=7, warm=2)
        print(f"FPS B={B} N=1024 G=512 pointnet2={pn2}: {us:.1f} us")
xyz = unit_sphere_clouds(2, 10000, torch.Generator().manual_seed(3)).to(dev)
for pn2 in (False, True):
    us = median_us(lambda: ua.fps_sample(xyz, 512, None, pointnet2=pn2), flush, n=7, warm=2)
    print(f"FPS B=2 N=10000 G=512 pointnet2={pn2} (cluster): {us:.1f} us")
