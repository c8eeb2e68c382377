"""DOTA.fit (ua_dota_fit_f32) at cfg 1 (K=40, D=512) and K=15, D=1024/1280: class-split (cluster size) sweep."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200 import _lib
dev = torch.device("cuda:0")
CFG = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
wbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev); rbuf = torch.ones(64 << 20, device=dev); sink = torch.zeros(1, device=dev)
for K, D in ((40, 512), (15, 1024), (40, 1024)):
    dota = ua.DOTA(CFG, D, K, torch.full((D, K), 0.001), device=dev)
    x1 = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
    y1 = torch.softmax(torch.randn(1, K, device=dev), 1)
    for ks in (1, 2, 4, 8):
        _lib.set_tuning("dota_ksplit", ks)
        for _ in range(3): dota.fit(x1, y1)
        ts = []
        for _ in range(12):
            wbuf.zero_(); sink.copy_(rbuf.sum())
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); dota.fit(x1, y1); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        us = sorted(ts)[6]
        print(f"K={K} D={D} ksplit={ks}: fit {us:.1f} us = {(8 * K * D * D + 4 * D * D) / us / 1e3:.0f} GB/s", flush=True)
    _lib.set_tuning("dota_ksplit", 0)
