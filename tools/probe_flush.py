"""How the L2 flush method changes the timed MODE-DOTA LVIS step (dirty lines left by a memset are written back inside
the timed kernel). Modes: w = 256 MiB memset; wr = memset then 256 MiB read; n = none (state 151 MB > 126 MB L2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200 import _lib
from uniadapter_b200.engine import MultiStreamModeDota
from oracle import synth

dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
wbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rbuf = torch.ones(64 << 20, dtype=torch.float32, device=dev)
sink = torch.zeros(1, device=dev)


def flush(mode):
    if mode in ("w", "wr"):
        wbuf.zero_()
    if mode == "wr":
        sink.copy_(rbuf.sum())


def timeit(fn, mode, n=15):
    ts = []
    for i in range(n + 3):
        flush(mode)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


S, K, M, D = 1, 1156, 8, 1024
text = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
model = MultiStreamModeDota(cfg, D, K, text, M, S, dev)
x = torch.nn.functional.normalize(torch.randn(S, 1, D, device=dev), dim=-1)
g = torch.softmax(100 * x @ text.t(), -1).contiguous()
for lp in (0, 1):
    _lib.set_tuning("modedota_logprod", lp)
    for mode in ("w", "wr", "n"):
        for name, call in (("pred+fit", lambda: model.step(x, x, g)), ("fit", lambda: model.step(None, x, g)),
                           ("pred", lambda: model.step(x, None, None))):
            med, mn = timeit(call, mode)
            by = (16 if "fit" in name else 8) * S * K * M * D
            print(f"lp={lp} flush={mode:7s} {name:8s}: {med:7.1f} / {mn:7.1f} us  {by / med / 1e3:8.1f} GB/s alg (median)", flush=True)
# reference point: a plain device copy of the same bytes under the same flush
a = torch.empty(K * M * D * 2, device=dev); b = torch.empty_like(a)
for mode in ("w", "wr", "n"):
    med, mn = timeit(lambda: b.copy_(a), mode)
    print(f"torch copy 75.8 MB -> 75.8 MB flush={mode:2s}: {med:7.1f} / {mn:7.1f} us  {2 * a.numel() * 4 / med / 1e3:8.1f} GB/s")
