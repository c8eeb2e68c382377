"""Per-stage CUDA-event breakdown of one engine step at the bench workload (eager, S streams)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200.encoders import build_encoder
from uniadapter_b200.engine import StreamEngine
from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
from uniadapter_b200.head import zero_shot_head

dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 15
CFG = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
enc = build_encoder('ulip', 0, dev)
text = synthetic_text_features(40, 512, 0)
eng = StreamEngine(enc, 'ulip', text, S, 1024, CFG, 8, True, dev, use_graph=False)
pc = unit_sphere_clouds(S, 1024, torch.Generator().manual_seed(1)).to(dev)
for _ in range(3):
    eng.step_device(pc)

def t(fn, n=5):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

with torch.no_grad():
    print(f"S={S}")
    print("tokenizer (Group)      ms", round(t(lambda: enc.point_encoder.group_divider(pc)), 3))
    nb, c = enc.point_encoder.group_divider(pc)
    print("mini-PointNet          ms", round(t(lambda: enc.point_encoder.encoder(nb)), 3))
    print("full encoder fwd       ms", round(t(lambda: enc(pc)), 3))
    f = enc(pc)
    print("head                   ms", round(t(lambda: zero_shot_head(f, eng.text)), 3))
print("adapt (2 enc + cache)  ms", round(t(eng._adapt), 3))
print("residual learning      ms", round(t(eng._learn_residuals), 3))
print("fuse                   ms", round(t(eng._fuse), 3))
print("whole eager step       ms", round(t(lambda: eng.step_device(pc)), 3))
