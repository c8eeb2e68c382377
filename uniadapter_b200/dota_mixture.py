"""MODE-DOTA: per-class diagonal Gaussian mixture cache with streaming-EM updates.

Host-side mirror of the reference's ``DOTA_mix`` (dota_mixture.py:7-274): same constructor, attributes
(``mu, var, pi, c, class_counts, t``) and methods (``fit, predict, update, _get_var, _log_likelihood``).
The state lives in HBM as contiguous (K,M,D) / (K,M) / (K,) fp32 tensors; ``predict`` / ``fit`` /
``predict_then_fit`` launch the fused sm_100a kernel ``ua_modedota_step_f32`` (csrc/modedota.cu).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


def init_state(clip_weights: torch.Tensor, num_modes: int, sigma_cfg: float, device):
    """Initial (mu, var, pi, c, class_counts, sigma_init) of dota_mixture.py:46-111, vectorised.

    mu[k,m] = text_k + e_{m mod D} * (0.1*sigma_init*(m+1)); var[:,m,:] = sigma_init*(1+0.05m) clamped at 1e-8;
    pi = c = 1/M; class_counts = 0. sigma_init = sigma if sigma < 0.1 else 1/D.
    """
    D, K = clip_weights.shape
    M = num_modes
    sigma_init = (1.0 / D) if sigma_cfg >= 0.1 else sigma_cfg
    clip_mu = clip_weights.t().to(device).float()
    mu = clip_mu.unsqueeze(1).repeat(1, M, 1).contiguous()
    delta_scale = sigma_init * 0.1
    for m in range(M):
        # same arithmetic as the reference: a zero offset row with one entry delta_scale*(m+1), added to the centre
        mu[:, m, m % D] = clip_mu[:, m % D] + torch.tensor(delta_scale * (m + 1), dtype=torch.float32)
    var = torch.ones(K, M, D, device=device) * sigma_init
    for m in range(M):
        var[:, m, :] *= (1.0 + 0.05 * m)
    var = torch.clamp(var, min=1e-8).contiguous()
    pi = (torch.ones(K, M, device=device) / M).contiguous()
    c = torch.full((K, M), 1.0 / M, device=device, dtype=torch.float32)
    class_counts = torch.zeros(K, device=device)
    return mu, var, pi, c, class_counts, sigma_init


class DOTA_mix(nn.Module):
    def __init__(self, cfg, input_shape, num_classes, clip_weights, num_modes=4, streaming_update_Sigma=True,
                 device=None):
        super().__init__()
        if device is None:
            device = 'cuda'
        self.device = torch.device(device)
        self.input_shape = input_shape
        self.num_classes = num_classes
        self.num_modes = num_modes
        if not streaming_update_Sigma:
            raise NotImplementedError("streaming_update_Sigma=False is not part of the hot path")
        self.streaming_update_Sigma = True
        self.epsilon = cfg.get('epsilon', 0.001)
        sigma_cfg = cfg.get('sigma', 1.0)
        if sigma_cfg >= 0.1:
            print(f"[DOTA-GMM] Warning: sigma={sigma_cfg} is too large for CLIP embeddings. "
                  f"Auto-corrected to 1/D = {1.0 / input_shape:.5f}")
        self.alpha_max = cfg.get('alpha_max', 0.5)
        if tuple(clip_weights.shape) != (input_shape, num_classes):
            raise ValueError(f"clip_weights must be (D,K)=({input_shape},{num_classes}), got {tuple(clip_weights.shape)}")
        (self.mu, self.var, self.pi, self.c, self.class_counts,
         self.sigma_init) = init_state(clip_weights, num_modes, sigma_cfg, self.device)
        self.t = 0

    # ---- reference private API (torch, differentiable: used by the residual text-alignment loss) ----------
    def _get_var(self):
        return torch.clamp(self.var + self.epsilon, min=1e-8)

    def _log_likelihood(self, x, mu, var):
        """(B,D),(K,M,D),(K,M,D) -> (B,K,M): -0.5*(sum log var + sum (x-mu)^2/var). Autograd path only
        (Uni_Adapter.py:219-228); the no-grad cache step never calls this, it runs in modedota.cu."""
        diff = x.unsqueeze(1).unsqueeze(2) - mu.unsqueeze(0)
        maha = torch.sum(diff ** 2 / var.unsqueeze(0), dim=-1)
        log_det = torch.sum(torch.log(var.unsqueeze(0)), dim=-1)
        return -0.5 * (log_det + maha)

    # ---- fused kernel entry ----------------------------------------------------------------------------------
    def _step(self, x_pred, x_fit, gamma_class):
        K, M, D = self.num_classes, self.num_modes, self.input_shape
        out = None
        Bp = B = 0
        if x_pred is not None:
            x_pred = x_pred.to(self.device).float().contiguous()
            Bp = x_pred.shape[0]
            out = torch.empty((Bp, K), dtype=torch.float32, device=self.device)
        if x_fit is not None:
            x_fit = x_fit.to(self.device).float().contiguous()
            gamma_class = gamma_class.to(self.device).float().contiguous()
            B = x_fit.shape[0]
            if tuple(gamma_class.shape) != (B, K):
                raise ValueError(f"gamma_class must be ({B},{K}), got {tuple(gamma_class.shape)}")
        rc = _lib.lib().ua_modedota_step_f32(
            _lib.ptr(x_pred), Bp, _lib.ptr(x_fit), _lib.ptr(gamma_class), B, K, 0, _lib.ptr(self.mu),
            _lib.ptr(self.var), _lib.ptr(self.pi), _lib.ptr(self.c), _lib.ptr(self.class_counts), 1, K, M, D,
            float(self.epsilon), _lib.ptr(out), K, 0, _lib.stream_ptr())
        _lib.check(rc, "ua_modedota_step_f32")
        if B:
            self.t += B
        return out

    # ---- reference public API ------------------------------------------------------------------------------------
    @torch.no_grad()
    def fit(self, x, gamma_class):
        self._step(None, x, gamma_class)

    @torch.no_grad()
    def predict(self, x, source_priors=None):
        scores = self._step(x, None, None)
        if source_priors is not None:
            p_est = self.class_counts / (self.class_counts.sum() + 1e-10)
            alpha_t = min(self.alpha_max, self.t / (self.t + 100.0))
            p_k = (1 - alpha_t) * source_priors.to(self.device) + alpha_t * p_est
            return scores + torch.log(p_k + 1e-10).unsqueeze(0)
        return scores

    @torch.no_grad()
    def predict_then_fit(self, x_pred, x_fit, gamma_class):
        """predict(x_pred) on the current state, then fit(x_fit, gamma_class) — one pass over the cache."""
        return self._step(x_pred, x_fit, gamma_class)

    @torch.no_grad()
    def sample_step(self, x, x_aug, prob_map):
        """The three cache operations of one batch-1 sample (Uni_Adapter.py:416-430) in ONE pass over the cache:
        ``predict(x.half())`` on the current state, ``fit(x, prob_map)``, ``fit(x_aug, prob_map)`` (``x_aug`` may be
        None). x, x_aug (1,D) normalised features, prob_map (1,K). Returns the cache logits (1,K); bit-identical to
        ``predict_then_fit(x.half(), x, prob_map)`` followed by ``fit(x_aug, prob_map)`` (csrc/modedota_sample.cu)."""
        K, M, D = self.num_classes, self.num_modes, self.input_shape
        x = x.to(self.device).float().contiguous()
        prob_map = prob_map.to(self.device).float().contiguous()
        if x.shape[0] != 1:
            raise ValueError("sample_step is the batch-1 step of the reference loop")
        x_aug = x_aug.to(self.device).float().contiguous() if x_aug is not None else None
        out = torch.empty((1, K), dtype=torch.float32, device=self.device)
        rc = _lib.lib().ua_modedota_sample_step_f32(
            _lib.ptr(x), _lib.ptr(x_aug), _lib.ptr(prob_map), K, 0, _lib.ptr(self.mu), _lib.ptr(self.var),
            _lib.ptr(self.pi), _lib.ptr(self.c), _lib.ptr(self.class_counts), 1, K, M, D, float(self.epsilon),
            _lib.ptr(out), K, 0, _lib.stream_ptr())
        if rc == _lib.UA_ERR_UNSUPPORTED:      # shape outside the register tiling: the same three operations, two passes
            out = self.predict_then_fit(x.half().float(), x, prob_map)
            if x_aug is not None:
                self.fit(x_aug, prob_map)
            return out
        _lib.check(rc, "ua_modedota_sample_step_f32")
        self.t += 2 if x_aug is not None else 1
        return out

    def update(self):
        """No-op (diagonal covariance needs no inversion); kept for API compatibility (dota_mixture.py:269-274)."""
        return None
