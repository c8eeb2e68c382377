import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import uniadapter_b200 as ua
from bench import L2Flush
from oracle import synth
from uniadapter_b200 import _lib
dev = torch.device("cuda:0")
flush = L2Flush(dev)
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
K, M, D = 1156, 8, 1024
text = torch.from_numpy(synth.unit_rows(K, D, 7)).to(dev)
x, xa, _ = synth.features(4, 1, D, text.cpu().numpy(), 8)
x, xa = torch.from_numpy(x).float().to(dev), torch.from_numpy(xa).float().to(dev)
full = ua.DOTA_mix(cfg, D, K, text.t().contiguous(), num_modes=M, device=dev)
prob = torch.softmax(100 * x[0] @ text.t(), 1)
xp = x[0].half().float()

def med(fn, n=11):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]

print("old predict+fit %.1f  old fit %.1f" % (med(lambda: full.predict_then_fit(xp, x[0], prob)), med(lambda: full.fit(xa[0], prob))))
for v, g in [(0, 0), (1, 1), (2, 1), (4, 1), (8, 1), (8, 2), (4, 2), (2, 2)]:
    _lib.set_tuning("sample_v", v); _lib.set_tuning("sample_g", g)
    try:
        a = med(lambda: full.sample_step(x[0], xa[0], prob)); b = med(lambda: full.sample_step(x[0], None, prob))
        print(f"sample_step V={v} G={g}: predict+fit+fit {a:.1f} us   predict+fit {b:.1f} us")
    except Exception as e:
        print(f"V={v} G={g}: {str(e)[:100]}")
