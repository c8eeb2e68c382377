"""Fusion of zero-shot logits with the cache logits (ua_fuse_logits_f32, csrc/fuse.cu).

mode 'mode_dota': Uni_Adapter.py:491-521 (entropy-weighted blend); mode 'dota': dota_mixture.py:289-293
(final = clip + w * dota). w = clamp(rho * c.mean() / batch, max=eta) is evaluated on device from ``c``.
"""
from __future__ import annotations

import torch

from . import _lib


def fuse_logits(clip_logits: torch.Tensor, dota_logits: torch.Tensor, c: torch.Tensor | None, rho: float, eta: float,
                batch: int, mode: str = 'mode_dota', *, c_sum: float | None = None, c_count: int | None = None,
                want_scaled: bool = False, per_row_c: bool = False):
    """Returns (final_logits (R,K), argmax (R,) int32, scaled_dota (R,K) | None).
    ``per_row_c``: ``c`` holds one adapter's counts per row (shape (R, ...)) instead of one shared adapter."""
    clip_logits = clip_logits.float().contiguous()
    R, K = clip_logits.shape
    is_f16 = dota_logits.dtype == torch.float16
    if not is_f16:
        dota_logits = dota_logits.float()
    if dota_logits.shape[0] == 1 and R > 1:
        dota_logits = dota_logits.expand(R, K)
    dota_logits = dota_logits.contiguous()
    dev = clip_logits.device
    out = torch.empty((R, K), dtype=torch.float32, device=dev)
    arg = torch.empty((R,), dtype=torch.int32, device=dev)
    scaled = torch.empty((R, K), dtype=torch.float32, device=dev) if want_scaled else None
    if c is not None:
        c = c.float().contiguous()
    count_c = int(c.numel()) if c is not None else 0
    stride = 0
    if per_row_c:
        count_c //= R
        stride = count_c
    total = float(c_count if c_count is not None else count_c)
    rc = _lib.lib().ua_fuse_logits_f32(_lib.ptr(clip_logits), _lib.ptr(dota_logits), int(is_f16), R, K, _lib.ptr(c),
                                       count_c, stride, float(c_sum) if c_sum is not None else -1.0, total, float(rho),
                                       float(eta), float(batch), 1 if mode == 'mode_dota' else 0, _lib.ptr(out),
                                       _lib.ptr(arg), _lib.ptr(scaled), _lib.stream_ptr())
    _lib.check(rc, "ua_fuse_logits_f32")
    return out, arg, scaled
