"""fp32-accurate tensor-core GEMM (csrc/gemm_tf32x3.cu): C = epilogue(A @ W^T) with A [M,K], W [N,K] (nn.Linear /
1x1-conv weight layout), operands carried as (hi, lo) tf32 pairs."""
from __future__ import annotations

import torch

from . import _lib


def split_tf32(x: torch.Tensor):
    """x (fp32, contiguous) -> (hi, lo): hi = tf32(x) with the low 13 mantissa bits zero, lo = x - hi."""
    x = x.float().contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    rc = _lib.lib().ua_split_tf32_f32(_lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(), _lib.stream_ptr())
    _lib.check(rc, "ua_split_tf32_f32")
    return hi, lo


def gemm_tf32x3(a, w, bias=None, group_bias=None, relu=False, out=False, out_split=False, group_max=False,
                group_max_split=False):
    """a = (a_hi, a_lo) [M,K]; w = (w_hi, w_lo) [N,K]. Returns a dict with the requested outputs:
    'out' [M,N] fp32, 'out_split' (hi, lo), 'gmax' [M/32,N], 'gmax_split' (hi, lo)."""
    a_hi, a_lo = a
    w_hi, w_lo = w
    M, K = a_hi.shape
    N = w_hi.shape[0]
    dev = a_hi.device
    res = {}
    o = torch.empty((M, N), dtype=torch.float32, device=dev) if out else None
    oh = torch.empty((M, N), dtype=torch.float32, device=dev) if out_split else None
    ol = torch.empty((M, N), dtype=torch.float32, device=dev) if out_split else None
    gm = torch.empty((M // 32, N), dtype=torch.float32, device=dev) if (group_max or group_max_split) else None
    gh = torch.empty_like(gm) if group_max_split else None
    gl = torch.empty_like(gm) if group_max_split else None
    rc = _lib.lib().ua_gemm_tf32x3_f32(
        _lib.ptr(a_hi), _lib.ptr(a_lo), a_hi.stride(0), _lib.ptr(w_hi), _lib.ptr(w_lo), w_hi.stride(0), M, N, K,
        _lib.ptr(bias), _lib.ptr(group_bias), int(bool(relu)), _lib.ptr(o), _lib.ptr(oh), _lib.ptr(ol), N, _lib.ptr(gm),
        _lib.ptr(gh), _lib.ptr(gl), _lib.stream_ptr())
    _lib.check(rc, "ua_gemm_tf32x3_f32")
    if out:
        res['out'] = o
    if out_split:
        res['out_split'] = (oh, ol)
    if gm is not None:
        res['gmax'] = gm
    if group_max_split:
        res['gmax_split'] = (gh, gl)
    return res
