"""One launch of every north-star kernel at its BASELINE shape, for `ncu --set full` (profiles/r2_*):
   python tools/prof_round2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uniadapter_b200 as ua
from oracle import synth
from uniadapter_b200 import parallel as PP
from uniadapter_b200.streams import unit_sphere_clouds
dev = torch.device("cuda:0")
ROUNDS = int(os.environ.get("PROF_ROUNDS", "2"))      # 1 under ncu (every matching launch is replayed ~40 times)
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rd = torch.ones(64 << 20, device=dev)


def cold():
    flush.zero_()
    rd.sum()


# cfg 4: LVIS cache pass (single-pass kernel; the round-1 predict+fit kernel beside it)
K, M, D = 1156, 8, 1024
text = torch.from_numpy(synth.unit_rows(K, D, 7)).to(dev)
x, xa, _ = synth.features(2, 1, D, text.cpu().numpy(), 8)
x, xa = torch.from_numpy(x).float().to(dev), torch.from_numpy(xa).float().to(dev)
full = ua.DOTA_mix(cfg, D, K, text.t().contiguous(), num_modes=M, device=dev)
prob = torch.softmax(100 * x[0] @ text.t(), 1)
for _ in range(ROUNDS):
    cold(); full.sample_step(x[0], xa[0], prob)
    cold(); full.predict_then_fit(x[0], x[0], prob)
# cfg 4 sharded step, ranks emulated on this GPU (P = 1: the product kernel with an empty exchange)
sh = PP.FusedShardedModeDota(cfg, text, M, dev, emulate_world=1, use_graph=False)
for _ in range(ROUNDS):
    cold(); sh.step(x[0] * 2.5, xa[0] * 1.5)
# cfg 1: DOTA fit + update + predict
Kd, Dd = 40, 512
a = ua.DOTA(cfg, Dd, Kd, torch.full((Dd, Kd), 0.001), device=dev)
xd = torch.nn.functional.normalize(torch.randn(1, Dd, device=dev), dim=-1)
yd = torch.softmax(torch.randn(1, Kd, device=dev), 1)
for _ in range(ROUNDS):
    cold(); a.fit(xd, yd)
    cold(); a.update()
    cold(); a.predict(xd.half())
# cfg 5 tokenizer: 64 clouds x 1024 coloured points, 512 groups x 64 neighbours; cfg 2 tokenizer: 30 clouds, k = 32
g = torch.Generator().manual_seed(5)
xyz = unit_sphere_clouds(64, 1024, g).to(dev)
rgb = torch.rand(64, 1024, 3, generator=g).to(dev)
for _ in range(ROUNDS):
    cold(); _, cen = ua.fps_sample(xyz, 512, None, pointnet2=True)
    cold(); ua.knn_group(xyz, cen, 64, rgb)
    cold(); _, cen2 = ua.fps_sample(xyz[:30], 512, torch.zeros(30, dtype=torch.long, device=dev))
    cold(); ua.knn_group(xyz[:30], cen2, 32)
# cfg 3 / 4 tokenizer: 2 clouds x 10 000 points (cluster FPS, ball query, kNN 64)
xyz10 = unit_sphere_clouds(2, 10000, g).to(dev)
pts10 = torch.cat((xyz10, torch.rand(2, 10000, 3, generator=g).to(dev)), -1).contiguous()
for _ in range(ROUNDS):
    cold(); _, c10 = ua.fps_sample(xyz10, 384, torch.zeros(2, dtype=torch.long, device=dev))
    cold(); ua.ball_group(xyz10, c10, 0.2, 64, pts10)
    cold(); _, c11 = ua.fps_sample(xyz10, 512, None, pointnet2=True)
    cold(); ua.knn_group(xyz10, c11, 64, pts10[..., 3:].contiguous())
# cfg 5 batched cache step
Kb, Bb = 55, 64
tb = torch.from_numpy(synth.unit_rows(Kb, D, 9)).to(dev)
mb = ua.DOTA_mix(cfg, D, Kb, tb.t().contiguous(), num_modes=M, device=dev)
xb = torch.nn.functional.normalize(torch.randn(Bb, D, device=dev), dim=-1)
gb = torch.softmax(100 * xb @ tb.t(), 1)
for _ in range(ROUNDS):
    cold(); mb.predict_then_fit(xb, xb, gb)
# cfg 2 cache step: 15 streams x 40 classes, M = 8, D = 512 (the kernel bench.py's headline roofline line quotes)
from uniadapter_b200.engine import MultiStreamModeDota
from uniadapter_b200.streams import synthetic_text_features
t2 = synthetic_text_features(40, 512, seed=1).to(dev)
c2 = MultiStreamModeDota(cfg, 512, 40, t2, 8, 15, dev)
x2 = torch.nn.functional.normalize(torch.randn(15, 512, device=dev), dim=-1)
xa2 = torch.nn.functional.normalize(x2 + 0.01 * torch.randn(15, 512, device=dev), dim=-1)
g2 = torch.softmax(100 * x2 @ t2.t(), 1).contiguous()
o2 = torch.zeros(15, 40, device=dev)
c2.sample_step(x2, xa2, g2, o2)
for _ in range(ROUNDS):
    cold(); c2.sample_step(x2, xa2, g2, o2)
# cfg 5 zero-shot head at batch 64 on the tcgen05 GEMM (HeadPlan: prepare + 3xTF32 GEMM + row statistics), K = 55 and 1156
xh = torch.randn(64, D, device=dev)
for Kh in (55, 1156):
    th = torch.from_numpy(synth.unit_rows(Kh, D, Kh)).to(dev)
    ua.zero_shot_head(xh, th)
    for _ in range(ROUNDS):
        cold(); ua.zero_shot_head(xh, th)
torch.cuda.synchronize()
print("ok")
