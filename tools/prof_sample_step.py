"""One LVIS-scale (K=1156, M=8, D=1024) single-pass cache step for ncu: python tools/prof_sample_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uniadapter_b200 as ua
from oracle import synth
dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
K, M, D = 1156, 8, 1024
text = torch.from_numpy(synth.unit_rows(K, D, 7)).to(dev)
x, xa, _ = synth.features(2, 1, D, text.cpu().numpy(), 8)
x, xa = torch.from_numpy(x).float().to(dev), torch.from_numpy(xa).float().to(dev)
full = ua.DOTA_mix(cfg, D, K, text.t().contiguous(), num_modes=M, device=dev)
prob = torch.softmax(100 * x[0] @ text.t(), 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(3):
    flush.zero_()
    full.sample_step(x[0], xa[0], prob)
    flush.zero_()
    full.predict_then_fit(x[0], x[0], prob)
torch.cuda.synchronize()
print("ok")
