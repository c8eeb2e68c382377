"""MODE-DOTA cache-step timing sweep (us, CUDA events, L2 flushed between launches) over the tuning knob modedota_v."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uniadapter_b200 as ua
from uniadapter_b200 import _lib
from uniadapter_b200.engine import MultiStreamModeDota
from oracle import synth

dev = torch.device("cuda:0")
cfg = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=15, do_flush=True):
    ts = []
    for i in range(n + 3):
        if do_flush:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for (S, K, M, D) in [(1, 1156, 8, 1024), (1, 145, 8, 1024), (15, 40, 8, 512), (1, 15, 8, 1280)]:
    text = torch.from_numpy(synth.unit_rows(K, D, 3)).to(dev)
    model = MultiStreamModeDota(cfg, D, K, text, M, S, dev)
    x = torch.nn.functional.normalize(torch.randn(S, 1, D, device=dev), dim=-1)
    g = torch.softmax(100 * x @ text.t(), -1).contiguous()
    for v, gr, lp in [(-1, 0, 0), (0, 0, 0), (0, 0, 1), (2, 1, 0), (4, 1, 0), (8, 1, 0), (2, 2, 0), (4, 2, 0), (8, 2, 0),
                      (4, 2, 1), (8, 2, 1), (8, 1, 1), (5, 1, 0), (5, 2, 0), (10, 2, 0), (10, 1, 0)]:
        if v > 0 and D % (128 * v):
            continue
        _lib.set_tuning("modedota_v", v)
        _lib.set_tuning("modedota_groups", gr)
        _lib.set_tuning("modedota_logprod", lp)
        for mode, call in (("pred+fit", lambda: model.step(x, x, g)), ("fit", lambda: model.step(None, x, g)),
                           ("pred", lambda: model.step(x, None, None))):
            n0 = _lib.launch_count()
            try:
                med, mn = timeit(call, do_flush=(K * M * D * 8 * S > 30e6))
            except Exception as ex:
                print("ERR", S, K, M, D, v, mode, ex); continue
            by = (16 if "fit" in mode else 8) * S * K * M * D
            print(f"S={S:2d} K={K:5d} M={M} D={D:5d} v={v:2d} g={gr} lp={lp} {mode:8s}: {med:8.1f} / {mn:8.1f} us  {by / mn / 1e3:8.1f} GB/s alg", flush=True)
    _lib.set_tuning("modedota_v", 0)
    _lib.set_tuning("modedota_logprod", 0)
    _lib.set_tuning("modedota_groups", 0)
