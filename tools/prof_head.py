"""Batched zero-shot head on the tcgen05 GEMM (cfg 5: B=64, D=1024, K=55 / 216 / 1156) for ncu: python tools/prof_head.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uniadapter_b200 as ua
from oracle import synth
dev = torch.device("cuda:0")
x = torch.randn(64, 1024, device=dev)
for K in (55, 1156):
    text = torch.from_numpy(synth.unit_rows(K, 1024, K)).to(dev)
    for _ in range(2):
        ua.zero_shot_head(x, text)            # B >= 64 -> HeadPlan (ua_head_prepare_f32 + ua_gemm_tf32x3_f32 + ua_row_stats_f32)
torch.cuda.synchronize()
print("ok")
