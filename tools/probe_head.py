"""Zero-shot head: SIMT kernels (ua_head_f32) vs the tcgen05 HeadPlan over batch sizes and class counts (device times, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uniadapter_b200 as ua
from uniadapter_b200.head import HeadPlan
from bench import L2Flush, median_us
dev = torch.device("cuda:0")
flush = L2Flush(dev)
D = 1024
for K in (55, 216, 1156):
    text = torch.nn.functional.normalize(torch.randn(K, D, generator=torch.Generator().manual_seed(K)), dim=-1).to(dev)
    plan = HeadPlan(text)
    for B in (8, 64, 256, 1024):
        x = torch.randn(B, D, device=dev)
        a = ua.zero_shot_head(x, text, tensor_cores=False)
        b = plan(x)
        err = float((a[1] - b[1]).abs().max())
        same_pred = bool(torch.equal(a[4], b[4]))
        t_simt = median_us(lambda: ua.zero_shot_head(x, text, tensor_cores=False), flush, n=9)
        t_tc = median_us(lambda: plan(x), flush, n=9)
        print(f"K={K:5d} B={B:5d}: SIMT {t_simt:7.1f} us   tcgen05 plan {t_tc:7.1f} us   max |logit diff| {err:.2e}  same argmax {same_pred}", flush=True)
