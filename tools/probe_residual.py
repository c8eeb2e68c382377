"""Timing of the residual-learning call (10 Adam steps): eager and CUDA-graph replay, over the tuning knobs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uniadapter_b200 import _lib
from uniadapter_b200.engine import MultiStreamModeDota
from uniadapter_b200.residual import ResidualLearner
from oracle import synth
dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
knobs = [(0, 0)] + [(c, d) for c in (1, 2, 3, 5) for d in (1, 2)]
for (S, K, M, D) in [(15, 40, 8, 512), (1, 40, 8, 512), (15, 15, 8, 1280), (8, 55, 8, 1024)]:
    text = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
    cache = MultiStreamModeDota(cfg, D, K, text, M, S, dev)
    x = torch.nn.functional.normalize(torch.randn(S, 1, D, device=dev), dim=-1)
    g = torch.softmax(100 * x @ text.t(), -1).contiguous()
    cache.step(None, x, g)
    for (cb, dbl) in (knobs if (S, K) == (15, 40) else knobs[:1]):
        _lib.set_tuning("resid_cb", cb); _lib.set_tuning("resid_dbl", dbl)
        try:
            learner = ResidualLearner(text, S, M, dev)
            for _ in range(2):
                learner.learn(cache.mu, cache.var, cache.pi, cache.epsilon)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                learner.learn(cache.mu, cache.var, cache.pi, cache.epsilon)
            gr.replay(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                gr.replay()
            e.record(); torch.cuda.synchronize()
            print(f"S={S} K={K} M={M} D={D} cb={cb} dbl={dbl}: graph replay {s.elapsed_time(e) / 10 * 1e3:.1f} us per learn() (10 Adam steps)", flush=True)
        except Exception as ex:
            print(f"S={S} K={K} cb={cb} dbl={dbl}: ERR {ex}")
    _lib.set_tuning("resid_cb", 0); _lib.set_tuning("resid_dbl", 0)
