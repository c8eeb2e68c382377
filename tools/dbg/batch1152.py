import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import uniadapter_b200 as ua
from uniadapter_b200 import _lib
from oracle import adapters as A, cases
dev = torch.device("cuda:0")
cfg = cases.CFG
for (S, K, M, D, Bp, B) in [(1, 6, 8, 1152, 2, 8), (1, 6, 8, 1024, 2, 8), (1, 6, 8, 768, 2, 8), (1, 6, 4, 1152, 2, 8)]:
    g = torch.Generator().manual_seed(1000 + K + D)
    text = torch.nn.functional.normalize(torch.randn(S, K, D, generator=g), dim=-1)
    lab = torch.randint(0, K, (S, B), generator=g)
    x_fit = torch.nn.functional.normalize(text[torch.arange(S)[:, None], lab] + 0.6 * torch.randn(S, B, D, generator=g) / D ** 0.5, dim=-1)
    x_pred = torch.nn.functional.normalize(torch.randn(S, max(Bp, 1), D, generator=g), dim=-1)
    gam = torch.softmax(100.0 * torch.einsum('sbd,skd->sbk', x_fit, text), -1)
    keys = ("mu", "var", "pi", "c", "class_counts")
    o = A.ModeDota(cfg, D, K, text[0].numpy().T, M); o2 = A.ModeDotaExactSum(cfg, D, K, text[0].numpy().T, M)
    for oo in (o, o2):
        for _ in range(2):
            oo.predict(x_pred[0].numpy()); oo.fit(x_fit[0].numpy(), gam[0].numpy())
    res = {}
    for mode in (0, -1):
        m = ua.DOTA_mix(cfg, D, K, text[0].t().contiguous().to(dev), num_modes=M, device=dev)
        st = {k_: getattr(m, k_).unsqueeze(0).contiguous() for k_ in keys}
        out = torch.zeros((S, max(Bp, 1), K), device=dev)
        _lib.set_tuning("modedota_batch", mode)
        for _ in range(2):
            rc = _lib.lib().ua_modedota_step_f32(_lib.ptr(x_pred.to(dev)), Bp, _lib.ptr(x_fit.to(dev)), _lib.ptr(gam.to(dev)), B, K, 0,
                 _lib.ptr(st["mu"]), _lib.ptr(st["var"]), _lib.ptr(st["pi"]), _lib.ptr(st["c"]), _lib.ptr(st["class_counts"]), S, K, M, D,
                 float(cfg['epsilon']), _lib.ptr(out), K, 0, _lib.stream_ptr())
            _lib.check(rc, "x")
        _lib.set_tuning("modedota_batch", 0)
        res[mode] = {k_: st[k_][0].cpu().numpy() for k_ in keys}
        print((S,K,M,D,Bp,B), "mode", mode, {k_: float(np.abs(res[mode][k_] - getattr(o, k_)).max()) for k_ in keys})
    print("  oracle twin:", {k_: float(np.abs(getattr(o2, k_) - getattr(o, k_)).max()) for k_ in keys})
    print("  batched vs general:", {k_: float(np.abs(res[0][k_] - res[-1][k_]).max()) for k_ in keys})
