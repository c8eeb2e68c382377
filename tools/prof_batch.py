"""ncu workload: batched MODE-DOTA cache step (cfg 5: K=216, M=8, D=1024, B=64) and the DOTA.update inverse (D=512)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200.engine import MultiStreamModeDota
from uniadapter_b200.streams import synthetic_text_features
dev = torch.device("cuda:0")
CFG = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
K, M, D, B = int(os.environ.get("PK", 216)), 8, 1024, 64
text = synthetic_text_features(K, D, 0).to(dev)
m = MultiStreamModeDota(CFG, D, K, text, M, 1, dev)
x = torch.nn.functional.normalize(torch.randn(1, B, D, device=dev), dim=-1)
g = torch.softmax(100 * x @ text.t(), -1).contiguous()
xp = x.mean(1, keepdim=True).contiguous()
for _ in range(3):
    m.step(xp, x, g)
dota = ua.DOTA(CFG, 512, 40, torch.full((512, 40), 0.001), device=dev)
x1 = torch.nn.functional.normalize(torch.randn(1, 512, device=dev), dim=-1)
y1 = torch.softmax(torch.randn(1, 40, device=dev), 1)
for _ in range(3):
    dota.fit(x1, y1); dota.update()
torch.cuda.synchronize()
# back-to-back launches: device time per launch without the Python launch gap
for name, fn, n in (("cache_step_b64", lambda: m.step(xp, x, g), 50), ("dota_update_d512", dota.update, 20)):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    print(name, round(s.elapsed_time(e) / n * 1e3, 1), "us per call (back to back, L2 warm)")
