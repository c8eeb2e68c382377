"""Kernel timing sweep on the GPU box (CUDA events, L2 flushed between timed launches). Prints a table; used to pick
tuning defaults and to fill DESIGN.md. Not part of the product."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua  # noqa: E402
from uniadapter_b200 import _lib  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    print("host cpus:", os.cpu_count(), "| gpu:", torch.cuda.get_device_name(0), "| torch", torch.__version__)
    print("tf32 matmul allowed:", torch.backends.cuda.matmul.allow_tf32)
    # ---- FPS -------------------------------------------------------------------------------------------
    print("\n== FPS (us, median/min) ==")
    for (B, N, G) in [(1, 1024, 512), (15, 1024, 512), (64, 1024, 512), (148, 1024, 512), (296, 1024, 512),
                      (1, 10000, 512), (8, 10000, 512), (1, 10000, 384), (1, 8192, 512)]:
        xyz = torch.from_numpy(synth.cloud(B, N, 1)).to(dev)
        for thr in ([0, 128, 256, 512, 1024] if N <= 2048 else [0, 512, 768, 1024]):
            _lib.set_tuning("fps_threads", thr)
            try:
                med, mn = timeit(lambda: ua.fps_sample(xyz, G, None, idx_dtype=torch.int32))
                print(f"fps B={B:4d} N={N:6d} G={G} threads={thr:5d}: {med:9.1f} / {mn:9.1f}  -> {B / med * 1e6:10.0f} clouds/s")
            except Exception as ex:
                print("fps", B, N, G, thr, "ERR", ex)
        _lib.set_tuning("fps_threads", 0)
    # ---- kNN / ball --------------------------------------------------------------------------------------
    print("\n== kNN group (us) ==")
    for (B, N, G, k) in [(1, 1024, 512, 32), (15, 1024, 512, 32), (64, 1024, 512, 64), (1, 10000, 512, 64), (8, 10000, 512, 64)]:
        xyz = torch.from_numpy(synth.cloud(B, N, 1)).to(dev)
        rgb = torch.rand(B, N, 3, device=dev)
        _, cen = ua.fps_sample(xyz, G, None)
        for w in [0, 1, 2, 4, 8]:
            _lib.set_tuning("knn_warps", w)
            med, mn = timeit(lambda: ua.knn_group(xyz, cen, k, rgb))
            print(f"knn B={B:3d} N={N:6d} G={G} k={k} warps={w}: {med:9.1f} / {mn:9.1f}")
        _lib.set_tuning("knn_warps", 0)
    xyz = torch.from_numpy(synth.cloud(1, 10000, 1)).to(dev)
    pts = torch.cat([xyz, torch.rand(1, 10000, 3, device=dev)], -1).contiguous()
    _, cen = ua.fps_sample(xyz, 384, None)
    med, mn = timeit(lambda: ua.ball_group(xyz, cen, 0.2, 64, pts))
    print(f"ball B=1 N=10000 S=384 ns=64: {med:9.1f} / {mn:9.1f}")
    # ---- head ----------------------------------------------------------------------------------------------
    print("\n== head (us) ==")
    for (B, D, K) in [(1, 512, 40), (1, 1024, 1156), (64, 1024, 55), (64, 1024, 1156), (15, 512, 40)]:
        x = torch.randn(B, D, device=dev)
        text = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
        med, mn = timeit(lambda: ua.zero_shot_head(x, text))
        by = 4 * (B * D + D * K + B * K)
        print(f"head B={B:3d} D={D} K={K:5d}: {med:8.1f} / {mn:8.1f}  ({by / mn / 1e3:8.1f} GB/s algorithmic)")
    # ---- MODE-DOTA -----------------------------------------------------------------------------------------
    print("\n== MODE-DOTA predict+fit (us) ==")
    cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
    for (K, M, D, B) in [(40, 8, 512, 1), (15, 8, 1280, 1), (1156, 8, 1024, 1), (55, 8, 1024, 64), (216, 8, 1024, 64),
                         (145, 8, 1024, 1), (289, 8, 1024, 1), (578, 8, 1024, 1)]:
        text = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
        model = ua.DOTA_mix(cfg, D, K, text.t().contiguous(), num_modes=M, device=dev)
        x = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=-1)
        g = torch.softmax(100 * x @ text.t(), 1)
        xp = x.mean(0, keepdim=True)
        for thr in [0, 256, 512, 1024]:
            _lib.set_tuning("modedota_threads", thr)
            try:
                med, mn = timeit(lambda: model.predict_then_fit(xp, x, g))
                by = 16 * K * M * D
                print(f"modedota K={K:5d} M={M} D={D} B={B:3d} thr={thr:4d}: {med:8.1f} / {mn:8.1f}  ({by / mn / 1e3:8.1f} GB/s algorithmic)")
            except Exception as ex:
                print("modedota", K, M, D, B, thr, "ERR", ex)
        _lib.set_tuning("modedota_threads", 0)
    # ---- DOTA ----------------------------------------------------------------------------------------------
    print("\n== DOTA (us) ==")
    for (K, D) in [(40, 512), (40, 1024)]:
        model = ua.DOTA(cfg, D, K, torch.full((D, K), 0.001), device=dev)
        x = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
        y = torch.softmax(torch.randn(1, K, device=dev), 1)
        med, mn = timeit(lambda: model.fit(x, y))
        by = 8 * K * D * D + 4 * D * D
        print(f"dota.fit K={K} D={D}: {med:8.1f} / {mn:8.1f} ({by / mn / 1e3:8.1f} GB/s algorithmic)")
        med, mn = timeit(lambda: model.update())
        print(f"dota.update (cuSOLVER inverse) D={D}: {med:8.1f} / {mn:8.1f}")
        med, mn = timeit(lambda: model.predict(x.half()))
        print(f"dota.predict K={K} D={D}: {med:8.1f} / {mn:8.1f}")
    print("\nlaunches so far:", _lib.launch_count())


if __name__ == "__main__":
    main()
