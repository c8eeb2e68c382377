"""Zero-shot cosine-logit head. Mirrors Uni_Adapter.py:21-26 (softmax_entropy) and :53-75 (get_logits_wrapper).

Text features follow one convention everywhere: ``text`` is (K,D) row-major with unit-norm rows and
``clip_weights = text.t()`` is the (D,K) matrix the reference multiplies with.
"""
from __future__ import annotations

import torch

from . import _lib


TENSOR_CORE_MIN_BATCH = 64      # rows from which the head is a dense contraction worth the tcgen05 GEMM (BASELINE cfg 5)


def zero_shot_head(x: torch.Tensor, text: torch.Tensor, scale: float = 100.0, *, want_prob: bool = True,
                   tensor_cores: bool | None = None):
    """x (B,D) raw encoder output, text (K,D) [or (S,K,D): one matrix per block of B/S rows]
    -> (xnorm, logits, entropy, prob, argmax int32 (B,)).

    Batch 1 (the reference loop) is a GEMV that streams the text matrix once: SIMT kernels. From
    ``TENSOR_CORE_MIN_BATCH`` rows on, with one shared text matrix and D % 32 == 0, the contraction runs on the tcgen05
    3xTF32 GEMM (:class:`HeadPlan`; the text rows are split per call, so a text matrix that residual learning keeps
    changing is always current). ``tensor_cores`` forces (True) or forbids (False) that path."""
    x = x.float().contiguous()
    text = text.float().contiguous()
    B, D = x.shape
    num_text = 1 if text.dim() == 2 else text.shape[0]
    K = text.shape[-2]
    if text.shape[-1] != D:
        raise ValueError(f"text features {tuple(text.shape)} do not match feature dim {D}")
    if tensor_cores is None:
        tensor_cores = B >= TENSOR_CORE_MIN_BATCH and num_text == 1 and D % 32 == 0 and x.is_cuda
    if tensor_cores:
        if num_text != 1 or D % 32:
            raise _lib.UaError("the tensor-core head needs one shared text matrix and D % 32 == 0")
        return HeadPlan(text.reshape(K, D), scale)(x, want_prob=want_prob)
    dev = x.device
    xnorm = torch.empty_like(x)
    logits = torch.empty((B, K), dtype=torch.float32, device=dev)
    prob = torch.empty((B, K), dtype=torch.float32, device=dev) if want_prob else None
    entropy = torch.empty((B,), dtype=torch.float32, device=dev)
    argmax = torch.empty((B,), dtype=torch.int32, device=dev)
    rc = _lib.lib().ua_head_f32(_lib.ptr(x), B, D, _lib.ptr(text), num_text, K, float(scale), _lib.ptr(xnorm), _lib.ptr(logits),
                                _lib.ptr(prob), _lib.ptr(entropy), _lib.ptr(argmax), _lib.stream_ptr())
    _lib.check(rc, "ua_head_f32")
    return xnorm, logits, entropy, prob, argmax


class HeadPlan:
    """The head as a dense contraction on the tcgen05 GEMM for batched inputs (B >= 64, BASELINE cfg 5): the text rows
    are padded to a multiple of 128 classes and split into (hi, lo) once; per call: normalise + scale + split the rows,
    one GEMM, softmax / entropy / argmax over the first K columns. Same outputs as :func:`zero_shot_head`."""

    def __init__(self, text: torch.Tensor, scale: float = 100.0):
        from .gemm import split_tf32
        text = text.float().contiguous()
        self.K, self.D = text.shape
        if self.D % 32:
            raise _lib.UaError(f"HeadPlan: feature dim {self.D} must be a multiple of 32")
        self.Kpad = -(-self.K // 128) * 128
        padded = torch.zeros(self.Kpad, self.D, dtype=torch.float32, device=text.device)
        padded[:self.K] = text
        self.text = split_tf32(padded)
        self.scale = float(scale)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, want_prob: bool = True):
        from .gemm import gemm_tf32x3
        x = x.float().contiguous()
        B, D = x.shape
        dev = x.device
        xnorm, hi, lo = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        rc = _lib.lib().ua_head_prepare_f32(_lib.ptr(x), B, D, self.scale, _lib.ptr(xnorm), _lib.ptr(hi), _lib.ptr(lo),
                                           _lib.stream_ptr())
        _lib.check(rc, "ua_head_prepare_f32")
        padded = gemm_tf32x3((hi, lo), self.text, out=True)['out']                      # (B, Kpad)
        prob = torch.empty((B, self.K), dtype=torch.float32, device=dev) if want_prob else None
        entropy = torch.empty((B,), dtype=torch.float32, device=dev)
        argmax = torch.empty((B,), dtype=torch.int32, device=dev)
        rc = _lib.lib().ua_row_stats_f32(_lib.ptr(padded), B, self.K, self.Kpad, _lib.ptr(prob), _lib.ptr(entropy),
                                        _lib.ptr(argmax), _lib.stream_ptr())
        _lib.check(rc, "ua_row_stats_f32")
        return xnorm, padded[:, :self.K], entropy, prob, argmax


def _as_text_rows(clip_weights: torch.Tensor, feat_dim: int) -> torch.Tensor:
    """Accept the reference's (D,K) ``clip_weights`` (or a (K,D) text matrix) and return (K,D) rows."""
    if clip_weights.shape[0] == feat_dim and clip_weights.shape[1] != feat_dim:
        return clip_weights.t().contiguous()
    if clip_weights.shape[1] == feat_dim and clip_weights.shape[0] != feat_dim:
        return clip_weights.contiguous()
    return clip_weights.t().contiguous()  # square: the reference convention is (D,K)


def encode(args, model, feature: torch.Tensor) -> torch.Tensor:
    """The encoder call of get_logits_wrapper (Uni_Adapter.py:54-66)."""
    if args.vlm3d == 'uni3d':
        return model.encode_pc(feature)
    xyz = feature[:, :, :3]
    if args.vlm3d == 'ulip':
        return model(xyz)
    if args.vlm3d == 'openshape':
        return model(xyz, feature)
    raise ValueError(f"unknown vlm3d {args.vlm3d!r}")


def get_logits_wrapper(args, model, feature: torch.Tensor, clip_weights: torch.Tensor):
    """Drop-in for Uni_Adapter.py:53-75: returns (pc_features, logits, loss, prob_map, pred).

    ``pred`` is a Python int for batch 1 (as the reference, which only supports batch 1 here — SURVEY D7) and an
    int32 tensor otherwise.
    """
    raw = encode(args, model, feature)
    text = _as_text_rows(clip_weights, raw.shape[-1])
    pc_features, logits, loss, prob_map, argmax = zero_shot_head(raw, text, 100.0)
    pred = int(argmax[0]) if argmax.numel() == 1 else argmax
    return pc_features, logits, loss, prob_map, pred


def softmax_entropy(x: torch.Tensor, enable_softmax: bool = True, temperature: float = 1.0) -> torch.Tensor:
    """Uni_Adapter.py:21-26 (host-side helper; the fused kernels compute the same quantity on device)."""
    probs = torch.softmax(x / temperature, dim=1) if enable_softmax else x
    return -(probs * torch.log(probs + 1e-10)).sum(dim=1)
