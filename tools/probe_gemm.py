"""Bring-up + timing of the tcgen05 3xTF32 GEMM against torch float64 / fp32 matmul."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uniadapter_b200.gemm import gemm_tf32x3, split_tf32
dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [(128, 128, 32), (128, 256, 32), (256, 128, 64), (128, 128, 128), (4096, 256, 128), (7695, 384, 384),
          (7680, 512, 256), (245760, 256, 128), (245760, 512, 256), (245760, 256, 512), (7695, 1152, 384), (7695, 1536, 384),
          (7695, 384, 1536)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
from uniadapter_b200 import _lib
bns = [int(b) for b in os.environ.get("GEMM_BNS", "0").split(",")]
for (M, N, K) in shapes:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    ap, wp = split_tf32(a), split_tf32(w)
    assert torch.equal(ap[0] + ap[1], a)
    r = gemm_tf32x3(ap, wp, bias=bias, out=True)['out']
    torch.cuda.synchronize()
    ref64 = (a.double() @ w.double().t() + bias.double())
    ref32 = a @ w.t() + bias
    err = (r.double() - ref64).abs().max().item()
    err32 = (ref32.double() - ref64).abs().max().item()
    scale = ref64.abs().max().item()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        gemm_tf32x3(ap, wp, bias=bias, out=True)
    s.record()
    for _ in range(5):
        gemm_tf32x3(ap, wp, bias=bias, out=True)
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 5 * 1e3
    for bn in bns:
        if bn and N % bn:
            continue
        _lib.set_tuning("gemm_bn", bn)
        for _ in range(2):
            gemm_tf32x3(ap, wp, bias=bias, out=True)
        s.record()
        for _ in range(5):
            gemm_tf32x3(ap, wp, bias=bias, out=True)
        e.record(); torch.cuda.synchronize()
        print(f"      bn={bn:3d}: {s.elapsed_time(e) / 5 * 1e3:8.1f} us")
    _lib.set_tuning("gemm_bn", 0)
    s.record()
    for _ in range(5):
        torch.addmm(bias, a, w.t())
    e.record(); torch.cuda.synchronize()
    us_t = s.elapsed_time(e) / 5 * 1e3
    print(f"M={M:6d} N={N:4d} K={K:4d}: max|err| {err:.3e} (torch fp32 {err32:.3e}, scale {scale:.2f})  "
          f"{us:8.1f} us = {2 * M * N * K / us / 1e6:7.1f} TFLOP/s fp32-equivalent | torch addmm {us_t:8.1f} us", flush=True)
    if M % 32 == 0:
        gb = torch.randn(M // 32, N, device=dev)
        res = gemm_tf32x3(ap, wp, bias=bias, group_bias=gb, relu=True, out=True, out_split=True, group_max_split=True)
        ref = torch.relu(ref64 + gb.double().repeat_interleave(32, 0)).float()
        e1 = (res['out'] - ref).abs().max().item()
        hi, lo = res['out_split']
        e2 = ((hi + lo) - res['out']).abs().max().item()
        e3 = (res['gmax'] - res['out'].view(M // 32, 32, N).amax(1)).abs().max().item()
        gh, gl = res['gmax_split']
        e4 = ((gh + gl) - res['gmax']).abs().max().item()
        lowbits = int((hi.view(torch.int32) & 0x1fff).abs().max())
        print(f"    epilogue: relu+group_bias err {e1:.3e}, split recombine {e2:.1e}, group max {e3:.1e}, gmax split {e4:.1e}, hi low bits {lowbits}")
