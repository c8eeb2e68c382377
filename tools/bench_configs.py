"""Per-config timing of the hot-path kernels at the five BASELINE.json shapes (1 GPU, CUDA events, L2 flushed between
launches for the stages whose working set exceeds a few MB). Prints one JSON line per config; tokenizer bytes are
SURVEY 8d's algorithmic figures (cloud in, centres + groups out)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200.engine import MultiStreamModeDota
from uniadapter_b200.head import HeadPlan
from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds

dev = torch.device("cuda:0")
CFG = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
wbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rbuf = torch.ones(64 << 20, device=dev)
sink = torch.zeros(1, device=dev)


def timed(fn, n=12, flush=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        if flush:
            wbuf.zero_(); sink.copy_(rbuf.sum())
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


def tokenizer(B, N, G, k, colored, ball=None):
    g = torch.Generator().manual_seed(B + N)
    xyz = unit_sphere_clouds(B, N, g).to(dev)
    rgb = torch.rand(B, N, 3, generator=g).to(dev) if colored else None
    start = torch.randint(0, N, (B,), generator=g).to(dev)
    out = {}
    out["fps_us"] = timed(lambda: ua.fps_sample(xyz, G, start))
    _, centers = ua.fps_sample(xyz, G, start)
    if ball is None:
        out["group_us"] = timed(lambda: ua.knn_group(xyz, centers, k, rgb))
        cout = 6 if colored else 3
    else:
        pts = torch.cat([xyz, rgb], -1).contiguous()
        out["group_us"] = timed(lambda: ua.ball_group(xyz, centers, ball, k, pts))
        cout = 9
    byts = B * (N * (6 if colored else 3) * 4 + G * 12 + G * k * cout * 4)
    tot = out["fps_us"] + out["group_us"]
    out.update(tokenizer_us=round(tot, 1), clouds_per_s=round(B / tot * 1e6), algorithmic_GBps=round(byts / tot / 1e3, 1))
    return {k_: (round(v, 1) if isinstance(v, float) else v) for k_, v in out.items()}


def cache(S, K, M, D, B):
    text = synthetic_text_features(K, D, 0).to(dev)
    m = MultiStreamModeDota(CFG, D, K, text, M, S, dev)
    x = torch.nn.functional.normalize(torch.randn(S, B, D, device=dev), dim=-1)
    g = torch.softmax(100 * x @ text.t(), -1).contiguous()
    xp = x.mean(1, keepdim=True).contiguous()
    us = timed(lambda: m.step(xp, x, g))
    return {"predict_fit_us": round(us, 1), "algorithmic_GBps": round(16 * S * K * M * D / us / 1e3, 1)}


def head(B, D, K):
    text = synthetic_text_features(K, D, 1).to(dev)
    x = torch.randn(B, D, device=dev)
    r = {"simt_us": round(timed(lambda: ua.zero_shot_head(x, text), flush=False), 1)}
    if B >= 64:
        plan = HeadPlan(text)
        r["tcgen05_us"] = round(timed(lambda: plan(x), flush=False), 1)
    return r


rows = []
# cfg 1: ULIP-2, DOTA (full covariance), batch 1
dota = ua.DOTA(CFG, 512, 40, torch.full((512, 40), 0.001), device=dev)
x1 = torch.nn.functional.normalize(torch.randn(1, 512, device=dev), dim=-1)
y1 = torch.softmax(torch.randn(1, 40, device=dev), 1)
rows.append({"config": 1, "what": "ULIP-2 1024 pts, 40 classes, DOTA", "tokenizer": tokenizer(1, 1024, 512, 32, False),
             "head": head(1, 512, 40),
             "dota": {"fit_us": round(timed(lambda: dota.fit(x1, y1)), 1), "fit_algorithmic_GBps": round((8 * 40 * 512 * 512 + 4 * 512 * 512) / timed(lambda: dota.fit(x1, y1)) / 1e3, 1),
                      "update_inverse_us": round(timed(lambda: dota.update(), flush=False), 1),
                      "predict_us": round(timed(lambda: dota.predict(x1.half()), flush=False), 1)}})
# cfg 2: 15 streams in lock-step (bench workload)
rows.append({"config": 2, "what": "ULIP-2 + MODE-DOTA M=8, 15 streams lock-step", "tokenizer": tokenizer(15, 1024, 512, 32, False),
             "head": head(15, 512, 40), "cache": cache(15, 40, 8, 512, 1)})
# cfg 3: OpenShape, 10k xyz+rgb points, ball query
rows.append({"config": 3, "what": "OpenShape 10k xyz+rgb pts, ball r=0.2 ns=64, 15 classes, MODE-DOTA M=8",
             "tokenizer": tokenizer(1, 10000, 384, 64, True, ball=0.2), "head": head(1, 1280, 15), "cache": cache(1, 15, 8, 1280, 1)})
# cfg 4: Uni3D-L, 10k points, LVIS cache
rows.append({"config": 4, "what": "Uni3D-L 10k pts (512 x 64), Objaverse-LVIS 1156-class MODE-DOTA M=8",
             "tokenizer": tokenizer(1, 10000, 512, 64, True), "head": head(1, 1024, 1156), "cache": cache(1, 1156, 8, 1024, 1)})
# cfg 5: batch 64 sweep
for K in (55, 216):
    rows.append({"config": 5, "what": f"Uni3D-L 1024 pts batch 64, {K} classes", "tokenizer": tokenizer(64, 1024, 512, 64, True),
                 "head": head(64, 1024, K), "cache": cache(1, K, 8, 1024, 64)})
for r in rows:
    print(json.dumps(r))
