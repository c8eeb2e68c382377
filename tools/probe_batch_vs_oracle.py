"""Batched (cluster) MODE-DOTA kernel and the general kernel against the CPU oracle on the same state (diagnostic)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200 import _lib
from oracle import adapters as A
dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
for (K, M, D, Bp, B) in ((15, 8, 1280, 32, 32), (15, 8, 1024, 32, 32), (9, 4, 512, 0, 16)):
    g = torch.Generator().manual_seed(1000 + K + D)
    text = torch.nn.functional.normalize(torch.randn(1, K, D, generator=g), dim=-1)
    lab = torch.randint(0, K, (1, B), generator=g)
    x_fit = torch.nn.functional.normalize(text[torch.arange(1)[:, None], lab] + 0.6 * torch.randn(1, B, D, generator=g) / D ** 0.5, dim=-1)
    x_pred = torch.nn.functional.normalize(torch.randn(1, max(Bp, 1), D, generator=g), dim=-1)
    gam = torch.softmax(100.0 * torch.einsum('sbd,skd->sbk', x_fit, text), -1)
    ora = [A.ModeDota(cfg, D, K, text[0].numpy().T, M), A.ModeDotaExactSum(cfg, D, K, text[0].numpy().T, M)]
    for o in ora:
        for _ in range(2):
            lo = o.predict(x_pred[0].numpy())
            o.fit(x_fit[0].numpy(), gam[0].numpy())
    res = {}
    for mode in (0, -1):
        m = ua.DOTA_mix(cfg, D, K, text[0].t().contiguous().to(dev), num_modes=M, device=dev)
        out = torch.zeros((1, max(Bp, 1), K), device=dev)
        X, XP, G_ = x_fit.to(dev).contiguous(), x_pred.to(dev).contiguous(), gam.to(dev).contiguous()
        _lib.set_tuning("modedota_batch", mode)
        for _ in range(2):
            rc = _lib.lib().ua_modedota_step_f32(_lib.ptr(XP) if Bp else None, Bp, _lib.ptr(X), _lib.ptr(G_), B, K, 0,
                                                 _lib.ptr(m.mu), _lib.ptr(m.var), _lib.ptr(m.pi), _lib.ptr(m.c),
                                                 _lib.ptr(m.class_counts), 1, K, M, D, float(cfg['epsilon']), _lib.ptr(out), K, 0,
                                                 _lib.stream_ptr())
            _lib.check(rc, "step")
        _lib.set_tuning("modedota_batch", 0)
        res[mode] = dict(mu=m.mu.cpu().numpy(), var=m.var.cpu().numpy(), c=m.c.cpu().numpy(), pi=m.pi.cpu().numpy(), lo=out[0].cpu().numpy())
    for mode, nm in ((0, "batched"), (-1, "general")):
        for oi, on in enumerate(("oracle", "exact-sum oracle")):
            o = ora[oi]
            d = {k: float(np.abs(res[mode][k] - getattr(o, k)).max()) for k in ("mu", "var", "c", "pi")}
            print(f"K={K} M={M} D={D} Bp={Bp} B={B} {nm:8s} vs {on:16s}", {k: f"{v:.2e}" for k, v in d.items()}, flush=True)
    print("   oracle vs exact-sum:", {k: f"{float(np.abs(getattr(ora[0], k) - getattr(ora[1], k)).max()):.2e}" for k in ("mu", "var", "c", "pi")})
    print("   batched vs general :", {k: f"{float(np.abs(res[0][k] - res[-1][k]).max()):.2e}" for k in ("mu", "var", "c", "pi", "lo")})
