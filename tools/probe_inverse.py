"""DOTA.update (dota.py:66-69) on B200: library inverses vs ua_dota_update_f32 (one cooperative launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uniadapter_b200 import _lib
dev = 'cuda'
for D in (128, 512, 768, 1024, 1280):
    g = torch.Generator().manual_seed(0)
    X = torch.randn(300, D, generator=g).to(dev) * 0.05
    S = (X.t() @ X) / 300 * 0.5 + 1e-4 * torch.eye(D, device=dev)
    S = ((S + S.t()) / 2).contiguous()
    eps = 1e-4
    A = (1 - eps) * S + eps * torch.eye(D, device=dev)
    I = torch.eye(D, device=dev)
    ws = torch.empty(_lib.lib().ua_dota_update_workspace_bytes(D), dtype=torch.uint8, device=dev)
    out_h = torch.empty((D, D), dtype=torch.float16, device=dev)
    out_f = torch.empty((D, D), dtype=torch.float32, device=dev)

    def ours():
        _lib.check(_lib.lib().ua_dota_update_f32(_lib.ptr(S), D, eps, _lib.ptr(ws), _lib.ptr(out_h), _lib.ptr(out_f),
                                                 _lib.stream_ptr()), "ua_dota_update_f32")
        return out_f

    def t(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n): r = fn()
        e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3, r
    ref = torch.linalg.inv(A.double())
    for name, fn in [("torch.inverse (reference, LU)", lambda: torch.linalg.inv(A)),
                     ("cholesky_ex + cholesky_solve(I)", lambda: torch.cholesky_solve(I, torch.linalg.cholesky_ex(A, check_errors=False)[0])),
                     ("ua_dota_update_f32", ours)]:
        us, r = t(fn)
        err = float((r.double() - ref).abs().max() / ref.abs().max())
        print(f"D={D:5d} {name:34s} {us:9.1f} us  rel err vs fp64 {err:.2e}  cond {float(torch.linalg.cond(A.double())):.1e}", flush=True)
