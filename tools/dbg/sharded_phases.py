import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench import L2Flush, median_us
from oracle import synth
from uniadapter_b200 import _lib, parallel as PP
dev = torch.device("cuda:0")
flush = L2Flush(dev)
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
K, M, D = 1156, 8, 1024
text = torch.from_numpy(synth.unit_rows(K, D, 7)).to(dev)
x, xa, _ = synth.features(2, 1, D, text.cpu().numpy(), 8)
x, xa = torch.from_numpy(x * 2.5).float().to(dev), torch.from_numpy(xa * 1.5).float().to(dev)
for P in (1, 8):
    sh = PP.FusedShardedModeDota(cfg, text, M, dev, emulate_world=P, use_graph=False)
    sh.step(x[0], xa[0]); sh.step(x[0], xa[0])
    for skip, what in [(0, "whole step"), (1, "without the class loop"), (3, "without class loop and fusion")]:
        _lib.set_tuning("sample_skip", skip)
        us = median_us(lambda: sh._launch(), flush, n=9)
        print(f"emulated P={P}: {what}: {us:.1f} us")
    _lib.set_tuning("sample_skip", 0)

import ctypes
for P in (1, 8):
    sh = PP.FusedShardedModeDota(cfg, text, M, dev, emulate_world=P, use_graph=False)
    _lib.set_tuning("sample_trace", 1)
    for _ in range(3):
        flush.zero_(); sh._launch(); torch.cuda.synchronize()
    buf = (ctypes.c_int64 * 16)()
    _lib.lib().ua_debug_sample_trace(buf)
    t = list(buf); t0 = t[0]
    names = {0: "start", 1: "rows normalised", 2: "own logits pushed", 3: "exchange 0 done + softmax stats", 4: "CTA 0 class loop done",
             8: "last CTA elected", 9: "exchange 1 done", 10: "cache row gathered", 11: "entropies", 12: "end"}
    print(f"emulated P={P} trace (us since start of rank 0's CTA 0):")
    for k in sorted(names):
        print(f"   {names[k]:34s} {(t[k] - t0) / 1e3:8.2f}")
    _lib.set_tuning("sample_trace", 0)
