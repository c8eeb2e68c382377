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
x, xa = torch.from_numpy(x * 2.5).float().to(dev), torch.from_numpy(xa * 1.5).float().to(dev)
full = ua.DOTA_mix(cfg, D, K, text.t().contiguous(), num_modes=M, device=dev)

def med(fn, n=11, cold=True):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        if cold: flush.zero_()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]

feats, clip_logits, _, prob, _ = ua.zero_shot_head(x[0], text)
feats_aug = ua.zero_shot_head(xa[0], text)[0]
for cold in (True, False):
    print("cold L2" if cold else "warm L2")
    print("  head (l2norm+logits+row_stats)      %.1f us" % med(lambda: ua.zero_shot_head(x[0], text), cold=cold))
    print("  sample_step (predict+fit+fit)        %.1f us" % med(lambda: full.sample_step(feats, feats_aug, prob), cold=cold))
    print("  sample_step (predict+fit, one fit)   %.1f us" % med(lambda: full.sample_step(feats, None, prob), cold=cold))
    print("  predict_then_fit (old kernel)        %.1f us" % med(lambda: full.predict_then_fit(feats.half().float(), feats, prob), cold=cold))
    print("  fit (old kernel)                     %.1f us" % med(lambda: full.fit(feats_aug, prob), cold=cold))
    dl = full.predict(feats)
    print("  predict only (old kernel)            %.1f us" % med(lambda: full.predict(feats), cold=cold))
    print("  fuse                                 %.1f us" % med(lambda: ua.fuse_logits(clip_logits, dl, full.c, 0.02, 0.1, 1, 'mode_dota'), cold=cold))
    src = torch.empty(2 * K * M * D, device=dev); dst = torch.empty_like(src)
    print("  plain copy of 151.5 MB               %.1f us" % med(lambda: dst.copy_(src), cold=cold))
    del src, dst
