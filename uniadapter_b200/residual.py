"""Residual text-feature learning (Uni_Adapter.py:191-270, 443-476): Adam on per-class text residuals against the
MODE-DOTA likelihood matrix.

The product path is :class:`ResidualLearner` / :func:`align_loss_grad`: the loss, its hand-derived backward and the
Adam update run in csrc/residual.cu (``ua_residual_learn_f32``), S streams per launch, no autograd.

The torch functions below remain as the reference-shaped API (``compute_text_alignment_loss`` is called by user code
with autograd) and as the plain-PyTorch fp32 reference the CUDA kernels are tested against. Two evaluations of the
same likelihood matrix LM[i,k] = logsumexp_m(log pi[k,m] + ll(x_i; k,m)):

* ``likelihood_matrix_broadcast`` — the reference's (K,K,M,D) broadcast through ``DOTA_mix._log_likelihood``
  (drop-in path, parity with the reference's own code);
* ``likelihood_matrix_gemm``      — the same quadratic form expanded into two GEMMs
  x^2 @ (1/v)^T and x @ (mu/v)^T (fp32 cuBLAS, batched over streams), which is what makes LVIS-scale and multi-stream
  residual learning fit in memory.
"""
from __future__ import annotations

import torch

from . import _lib


def likelihood_matrix_broadcast(class_embeddings, model):
    log_lik = model._log_likelihood(class_embeddings, model.mu, model._get_var())          # (K,K,M)
    return torch.logsumexp(torch.log(model.pi + 1e-10).unsqueeze(0) + log_lik, dim=2)       # (K,K)


def gemm_operands(mu, var, pi, eps):
    """Per-(stream,class,mode) operands of the expanded form; mu,var (...,K,M,D), pi (...,K,M)."""
    v = torch.clamp(var + eps, min=1e-8)
    iv = 1.0 / v
    a = iv.flatten(-3, -2)                                   # (...,K*M,D)
    b = (mu * iv).flatten(-3, -2)
    const = ((mu * mu * iv).sum(-1) + torch.log(v).sum(-1)).flatten(-2, -1)   # (...,K*M)
    return a, b, const, torch.log(pi + 1e-10)


def likelihood_matrix_gemm(class_embeddings, operands, K, M):
    """class_embeddings (...,K,D) -> LM (...,K,K)."""
    a, b, const, log_pi = operands
    x = class_embeddings
    maha = (x * x) @ a.transpose(-1, -2) - 2.0 * (x @ b.transpose(-1, -2)) + const.unsqueeze(-2)   # (...,K,K*M)
    ll = (-0.5 * maha).unflatten(-1, (K, M))
    return torch.logsumexp(log_pi.unsqueeze(-3) + ll, dim=-1)


def alignment_loss_from_matrix(lm):
    """Uni_Adapter.py:245-251: double-exp normalised contrastive loss of a (...,K,K) likelihood matrix."""
    e = torch.exp(torch.exp(lm / lm.amax(dim=(-2, -1), keepdim=True)))
    diag = torch.diagonal(e, dim1=-2, dim2=-1)
    return -(diag / e.sum(dim=-1)).mean(-1) - (diag / e.sum(dim=-2)).mean(-1)


def compute_text_alignment_loss(class_embeddings, mode_dota_model):
    """Drop-in for Uni_Adapter.py:191-270: returns (loss, likelihood_matrix)."""
    if not class_embeddings.requires_grad:
        raise RuntimeError("class_embeddings must require gradients for optimization")
    lm = likelihood_matrix_broadcast(class_embeddings, mode_dota_model)
    return alignment_loss_from_matrix(lm), lm


# ----------------------------------------------------------------------------------------------------------
# CUDA path (csrc/residual.cu)
# ----------------------------------------------------------------------------------------------------------
def _state_dims(mu):
    if mu.dim() == 3:
        return (1,) + tuple(mu.shape)
    return tuple(mu.shape)


def _scratch(S, K, M, D, device):
    n = int(_lib.lib().ua_residual_scratch_floats(S, K, M, D))
    if n < 0:
        raise _lib.UaError(f"residual learning does not support S={S} K={K} M={M} D={D}: "
                           + _lib.lib().ua_last_error().decode())
    return torch.empty(n, dtype=torch.float32, device=device)


@torch.no_grad()
def align_loss_grad(text0, residual, mu, var, pi, eps, want_grad=True):
    """One evaluation of the alignment loss on the device: text0 (K,D)|(S,K,D), residual (S,K,D)|(K,D),
    mu,var (S,K,M,D)|(K,M,D), pi (S,K,M)|(K,M) -> loss (S,), likelihood matrix (S,K,K), d loss / d residual (S,K,D),
    normalised embeddings (S,K,D)."""
    S, K, M, D = _state_dims(mu)
    dev = mu.device
    text0 = text0.to(dev).float().contiguous()
    residual = residual.to(dev).float().contiguous()
    mu, var, pi = mu.float().contiguous(), var.float().contiguous(), pi.float().contiguous()
    stride = K * D if text0.dim() == 3 else 0
    emb = torch.empty((S, K, D), dtype=torch.float32, device=dev)
    loss = torch.empty((S,), dtype=torch.float32, device=dev)
    lm = torch.empty((S, K, K), dtype=torch.float32, device=dev)
    grad = torch.empty((S, K, D), dtype=torch.float32, device=dev) if want_grad else None
    scratch = _scratch(S, K, M, D, dev)
    rc = _lib.lib().ua_align_loss_grad_f32(_lib.ptr(text0), stride, _lib.ptr(residual), _lib.ptr(mu), _lib.ptr(var),
                                          _lib.ptr(pi), S, K, M, D, float(eps), _lib.ptr(emb), _lib.ptr(loss),
                                          _lib.ptr(lm), _lib.ptr(grad), _lib.ptr(scratch), scratch.numel(),
                                          _lib.stream_ptr())
    _lib.check(rc, "ua_align_loss_grad_f32")
    return loss, lm, grad, emb


class ResidualLearner:
    """Per-stream text residuals with their Adam state (torch.optim.Adam defaults, Uni_Adapter.py:346-352), advanced
    by ``learn`` — 10 zero_grad / backward / step rounds of the reference in one library call. All buffers are static,
    the Adam step counter lives on the device: ``learn`` can be captured into a CUDA graph."""

    def __init__(self, text0, num_streams, num_modes, device, lr=1e-3, betas=(0.9, 0.999), adam_eps=1e-8):
        self.dev = torch.device(device)
        self.text0 = text0.to(self.dev).float().contiguous()
        self.S, self.M = num_streams, num_modes
        self.K, self.D = self.text0.shape[-2:]
        self.lr, self.betas, self.adam_eps = lr, betas, adam_eps
        S, K, D = self.S, self.K, self.D
        self.residual = torch.zeros(S, K, D, device=self.dev)
        self.adam_m = torch.zeros(S, K, D, device=self.dev)
        self.adam_v = torch.zeros(S, K, D, device=self.dev)
        self.adam_t = torch.zeros(S, dtype=torch.int32, device=self.dev)
        self.text = torch.empty(S, K, D, device=self.dev)      # normalize(text0 + residual): the head's text matrix
        self.scratch = _scratch(S, K, num_modes, D, self.dev)
        self.learn(None, None, None, 0.0, iters=0)

    @torch.no_grad()
    def learn(self, mu, var, pi, eps, iters=10, loss_out=None):
        """iters Adam steps on the alignment loss of the cache (mu,var (S,K,M,D), pi (S,K,M)); refreshes ``text``."""
        S, K, M, D = self.S, self.K, self.M, self.D
        if iters > 0:
            if tuple(mu.shape[-3:]) != (K, M, D) or mu.numel() != S * K * M * D:
                raise ValueError(f"cache state must be ({S},{K},{M},{D}), got {tuple(mu.shape)}")
            mu_p, var_p, pi_p = _lib.ptr(mu), _lib.ptr(var), _lib.ptr(pi)
        else:   # text refresh only: the cache is not read
            mu_p = var_p = pi_p = _lib.ptr(self.scratch)
        stride = K * D if self.text0.dim() == 3 else 0
        rc = _lib.lib().ua_residual_learn_f32(
            _lib.ptr(self.text0), stride, _lib.ptr(self.residual), _lib.ptr(self.adam_m), _lib.ptr(self.adam_v),
            _lib.ptr(self.adam_t), mu_p, var_p, pi_p, S, K, M, D, float(eps), float(self.lr), float(self.betas[0]),
            float(self.betas[1]), float(self.adam_eps), int(iters), _lib.ptr(self.text), _lib.ptr(loss_out),
            _lib.ptr(self.scratch), self.scratch.numel(), _lib.stream_ptr())
        _lib.check(rc, "ua_residual_learn_f32")
        return self.text
