"""Bring-up + timing of the tcgen05 attention kernel against torch SDPA (float64 reference)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from uniadapter_b200.gemm import attention_tf32x3
dev = torch.device("cuda:0")
torch.manual_seed(0)
cases = [(1, 128, 1), (1, 130, 2), (2, 513, 6), (15, 513, 6)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for (B, N, H) in cases:
    C = H * 64
    qkv = torch.randn(B * N, 3 * C, device=dev)
    hi, lo = attention_tf32x3(qkv, B, N, H)
    torch.cuda.synchronize()
    out = (hi + lo).view(B, N, C)
    q, k, v = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double()).transpose(1, 2).reshape(B, N, C)
    ref32 = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C)
    err = float((out.double() - ref).abs().max()); err32 = float((ref32.double() - ref).abs().max())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        attention_tf32x3(qkv, B, N, H)
    s.record()
    for _ in range(5):
        attention_tf32x3(qkv, B, N, H)
    e.record(); torch.cuda.synchronize(); us = s.elapsed_time(e) / 5 * 1e3
    from uniadapter_b200 import _lib
    tl = _lib.enable_kernel_timing(True)
    for _ in range(5):
        attention_tf32x3(qkv, B, N, H)
    torch.cuda.synchronize()
    print("   ", {k_: round(v_[2], 1) for k_, v_ in tl.summary().items()})
    _lib.enable_kernel_timing(False)
    s.record()
    for _ in range(5):
        F.scaled_dot_product_attention(q, k, v)
    e.record(); torch.cuda.synchronize(); us_t = s.elapsed_time(e) / 5 * 1e3
    print(f"B={B} N={N} H={H}: max|err| {err:.3e} (torch fp32 SDPA {err32:.3e}, max|ref| {float(ref.abs().max()):.2f})  "
          f"{us:8.1f} us (prepare + attention) | torch SDPA {us_t:8.1f} us", flush=True)
