"""Timing of the residual-learning call (10 Adam steps) at the bench shape: CUDA path vs the torch autograd path."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200 import _lib
from uniadapter_b200.engine import MultiStreamModeDota
from uniadapter_b200.residual import ResidualLearner
from oracle import synth
dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
for (S, K, M, D) in [(15, 40, 8, 512), (1, 40, 8, 512), (15, 15, 8, 1280), (8, 55, 8, 1024)]:
    text = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
    cache = MultiStreamModeDota(cfg, D, K, text, M, S, dev)
    x = torch.nn.functional.normalize(torch.randn(S, 1, D, device=dev), dim=-1)
    g = torch.softmax(100 * x @ text.t(), -1).contiguous()
    cache.step(None, x, g)
    learner = ResidualLearner(text, S, M, dev)
    for _ in range(3):
        learner.learn(cache.mu, cache.var, cache.pi, cache.epsilon)
    torch.cuda.synchronize()
    tl = _lib.enable_kernel_timing(True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        learner.learn(cache.mu, cache.var, cache.pi, cache.epsilon)
    e.record(); torch.cuda.synchronize()
    _lib.enable_kernel_timing(False)
    print(f"S={S} K={K} M={M} D={D}: eager {s.elapsed_time(e) / 5 * 1e3:.1f} us per learn() (10 Adam steps)")
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        learner.learn(cache.mu, cache.var, cache.pi, cache.epsilon)
    gr.replay(); torch.cuda.synchronize()
    s.record()
    for _ in range(10):
        gr.replay()
    e.record(); torch.cuda.synchronize()
    print(f"    CUDA graph replay {s.elapsed_time(e) / 10 * 1e3:.1f} us per learn()")
