"""ncu workload for the round summary: one launch set of every hot kernel at the bench / cfg-4 shapes.
  python tools/prof_all.py            (then: ncu --set full -k regex:... python tools/prof_all.py)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200.encoders import build_encoder
from uniadapter_b200.engine import MultiStreamModeDota
from uniadapter_b200.residual import ResidualLearner
from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
from oracle import synth

dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
S = 15
# 1. encoder forward at the bench shape: tokenizer (fps, knn), group encoder + Linear GEMMs
from uniadapter_b200.encoders import UlipPointBert, use_tensor_cores
torch.manual_seed(0)
enc = use_tensor_cores(UlipPointBert(depth=1).to(dev).eval())      # one block: every GEMM shape of the step once
pc = unit_sphere_clouds(S, 1024, torch.Generator().manual_seed(1)).to(dev)
with torch.no_grad():
    feats = enc(pc)
# 2. head + cache step + residual learning at the bench shape
text = synthetic_text_features(40, 512, 0).to(dev)
xn, logits, _, prob, _ = ua.zero_shot_head(feats, text)
cache = MultiStreamModeDota(cfg, 512, 40, text, 8, S, dev)
x = xn.unsqueeze(1).contiguous()
cache.step(x, x, prob.unsqueeze(1).contiguous())
learner = ResidualLearner(text, S, 8, dev)
learner.learn(cache.mu, cache.var, cache.pi, cache.epsilon, iters=1)
# 3. the LVIS-scale cache step (cfg 4)
K, M, D = 1156, 8, 1024
t4 = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
model = ua.DOTA_mix(cfg, D, K, t4.t().contiguous(), num_modes=M, device=dev)
x4 = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
g4 = torch.softmax(100 * x4 @ t4.t(), 1)
model.predict_then_fit(x4, x4, g4)
torch.cuda.synchronize()
print("ok")
