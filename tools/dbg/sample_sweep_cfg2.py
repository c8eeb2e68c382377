import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench import L2Flush, median_us, CFG
from uniadapter_b200 import _lib
from uniadapter_b200.engine import MultiStreamModeDota
from uniadapter_b200.streams import synthetic_text_features
dev = torch.device("cuda:0")
flush = L2Flush(dev)
for (S, K, M, D) in [(15, 40, 8, 512), (4, 15, 8, 1280), (1, 40, 8, 512)]:
    text = synthetic_text_features(K, D, seed=1).to(dev)
    cache = MultiStreamModeDota(CFG, D, K, text, M, S, dev)
    x = torch.nn.functional.normalize(torch.randn(S, D, device=dev), dim=-1)
    xa = torch.nn.functional.normalize(x + 0.01 * torch.randn(S, D, device=dev), dim=-1)
    g = torch.softmax(100 * x @ text.t(), 1).contiguous()
    out = torch.zeros(S, K, device=dev)
    for v, gg in [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 1), (2, 2), (4, 1), (4, 2), (5, 1), (5, 2), (10, 1), (10, 2)]:
        _lib.set_tuning("sample_v", v); _lib.set_tuning("sample_g", gg)
        try:
            ok = cache.sample_step(x, xa, g, out)
            if not ok:
                continue
            us = median_us(lambda: cache.sample_step(x, xa, g, out), flush, n=9)
            print(f"S={S} K={K} M={M} D={D} V={v} G={gg}: {us:.1f} us")
        except Exception as e:
            print(f"V={v} G={gg}: {str(e)[:80]}")
    _lib.set_tuning("sample_v", 0); _lib.set_tuning("sample_g", 0)
    for gg in (1, 2):
        for per in (1, 2, 3, 4):
            _lib.set_tuning("sample_g", gg); _lib.set_tuning("sample_per", per)
            try:
                us = median_us(lambda: cache.sample_step(x, xa, g, out), flush, n=9)
                print(f"S={S} K={K} M={M} D={D} G={gg} CTAs/SM={per}: {us:.1f} us")
            except Exception as e:
                print(f"G={gg} per={per}: {str(e)[:80]}")
    _lib.set_tuning("sample_g", 0); _lib.set_tuning("sample_per", 0)
