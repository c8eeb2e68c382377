"""Round-2 (second session) probes, one process: python tools/probe_r2b.py [knn] [dota] [cache] ...
  knn  : register-mask kNN selection vs the candidate-buffer histogram path vs the streaming filter (bit equality + times)
  dota : staged DOTA fit, loads-in-flight sweep x programmatic dependent launch of the mean kernel (equality + times)
Device times: CUDA events, L2 flushed (memset + read), median."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uniadapter_b200 as ua
from uniadapter_b200 import _lib
from bench import L2Flush
from uniadapter_b200.streams import unit_sphere_clouds

dev = torch.device("cuda:0")
flush = L2Flush(dev)
what = set(sys.argv[1:]) or {"knn", "dota"}


def med(fn, n=9, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


if "knn" in what:
    print("== kNN grouping: knn_hist 0 = register masks (new), 2 = candidate-buffer histogram, -1 = streaming filter", flush=True)
    cases = [(64, 1024, 512, 64, True, "plain"), (15, 1024, 512, 32, False, "plain"), (5, 700, 128, 32, True, "plain"),
             (4, 1024, 256, 64, True, "dup"), (4, 1024, 256, 32, False, "grid"), (3, 516, 64, 8, False, "plain"),
             (2, 1024, 64, 128, True, "plain"), (2, 100, 10, 100, False, "plain"), (2, 1024, 64, 1, False, "plain")]
    for B, N, G, k, col, kind in cases:
        g = torch.Generator().manual_seed(B * 1000 + N + k)
        xyz = unit_sphere_clouds(B, N, g)
        if kind == "dup":
            xyz[:, N // 2:] = xyz[:, : N - N // 2]
        if kind == "grid":
            xyz = torch.round(xyz * 8) / 8
        xyz = xyz.to(dev)
        rgb = torch.rand(B, N, 3, generator=g).to(dev) if col else None
        _, centers = ua.fps_sample(xyz, G, None)
        outs = {}
        for mode in (0, 2, -1):
            _lib.set_tuning("knn_hist", mode)
            idx, neigh, feat = ua.knn_group(xyz, centers, k, rgb, want_idx=True)
            torch.cuda.synchronize()
            outs[mode] = (idx, neigh, feat)
        _lib.set_tuning("knn_hist", 0)
        ok = all(torch.equal(outs[0][i], outs[-1][i]) for i in range(3) if outs[0][i] is not None)
        ok2 = all(torch.equal(outs[2][i], outs[-1][i]) for i in range(3) if outs[2][i] is not None)
        print(f"  B={B} N={N} G={G} k={k} colour={col} {kind}: masks == streaming: {ok}; histogram == streaming: {ok2}", flush=True)
    N, G, k = 1024, 512, 64
    for B in (15, 64, 148, 592, 1184):
        g = torch.Generator().manual_seed(B)
        xyz = unit_sphere_clouds(B, N, g).to(dev)
        rgb = torch.rand(B, N, 3, generator=g).to(dev)
        _, centers = ua.fps_sample(xyz, G, None)
        row = []
        for mode, w in ((0, 0), (2, 0), (-1, 0)):
            _lib.set_tuning("knn_hist", mode)
            _lib.set_tuning("knn_warps", w)
            t64 = med(lambda: ua.knn_group(xyz, centers, 64, rgb))
            t32 = med(lambda: ua.knn_group(xyz, centers, 32))
            row.append(f"hist={mode} W={w}: k64+rgb {t64:7.1f} k32 {t32:7.1f}")
        _lib.set_tuning("knn_hist", 0)
        _lib.set_tuning("knn_warps", 0)
        f = med(lambda: ua.fps_sample(xyz, G, None))
        print(f"  B={B:5d} fps {f:7.1f} us | " + " | ".join(row), flush=True)

if "dota" in what:
    print("== DOTA fit (ua_dota_fit_f32), batch 1", flush=True)
    CFG = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
    for K, D in ((40, 512), (15, 1280), (40, 1024)):
        x1 = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
        y1 = torch.softmax(torch.randn(1, K, device=dev), 1)

        def state_after(n, **tune):
            for kk, v in tune.items():
                _lib.set_tuning(kk, v)
            torch.manual_seed(0)
            d = ua.DOTA(CFG, D, K, torch.full((D, K), 0.001), device=dev)
            for _ in range(n):
                d.fit(x1, y1)
            torch.cuda.synchronize()
            return d

        ref = state_after(3, dota_staged=0)
        _lib.set_tuning("dota_staged", 1)
        for ka in (8, 10, 20):
            for pdl in (0, 1):
                d = state_after(3, dota_ka=ka, dota_pdl=pdl)
                same = all(torch.equal(getattr(d, a), getattr(ref, a)) for a in ("mu", "c", "Sigma", "overall_Sigma") if hasattr(d, a))
                close = max(float((getattr(d, a) - getattr(ref, a)).abs().max()) for a in ("mu", "c", "Sigma", "overall_Sigma") if hasattr(d, a))
                us = med(lambda: d.fit(x1, y1), n=15, warm=3)
                by = 8 * K * D * D + 4 * D * D
                print(f"  K={K} D={D} ka={ka} pdl={pdl}: {us:6.1f} us = {by / us / 1e3:6.0f} GB/s   bit-equal to the general kernel: {same} (max diff {close:.2e})", flush=True)
        _lib.set_tuning("dota_ka", 0)
        _lib.set_tuning("dota_pdl", 1)
