"""ncu workload: the LVIS-scale MODE-DOTA predict+fit (cfg 4), a few launches. UA_TUNING selects the variant."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from oracle import synth
dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
K, M, D = 1156, 8, 1024
text = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
model = ua.DOTA_mix(cfg, D, K, text.t().contiguous(), num_modes=M, device=dev)
x = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
g = torch.softmax(100 * x @ text.t(), 1)
for _ in range(5):
    model.predict_then_fit(x, x, g)
torch.cuda.synchronize()
print("ok")
