"""Residual text-feature learning (Uni_Adapter.py:191-270, 443-476): Adam on per-class text residuals against the
MODE-DOTA likelihood matrix. Stays in PyTorch autograd (kept, row §8f-1 is its CUDA rewrite).

Two evaluations of the same likelihood matrix LM[i,k] = logsumexp_m(log pi[k,m] + ll(x_i; k,m)):

* ``likelihood_matrix_broadcast`` — the reference's (K,K,M,D) broadcast through ``DOTA_mix._log_likelihood``
  (drop-in path, parity with the reference's own code);
* ``likelihood_matrix_gemm``      — the same quadratic form expanded into two GEMMs
  x^2 @ (1/v)^T and x @ (mu/v)^T (fp32 cuBLAS, batched over streams), which is what makes LVIS-scale and multi-stream
  residual learning fit in memory.
"""
from __future__ import annotations

import torch


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
