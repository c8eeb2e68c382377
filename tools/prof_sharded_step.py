"""Class-sharded sample step at LVIS scale for ncu, ranks emulated on one GPU: python tools/prof_sharded_step.py [P]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from uniadapter_b200 import parallel as PP
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
K, M, D = 1156, 8, 1024
text = torch.from_numpy(synth.unit_rows(K, D, 7)).to(dev)
x, xa, _ = synth.features(2, 1, D, text.cpu().numpy(), 8)
x, xa = torch.from_numpy(x * 2.5).float().to(dev), torch.from_numpy(xa * 1.5).float().to(dev)
sh = PP.FusedShardedModeDota(cfg, text, M, dev, emulate_world=P, use_graph=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(4):
    flush.zero_()
    sh.step(x[0], xa[0])
torch.cuda.synchronize()
sh.check()
print("ok")
