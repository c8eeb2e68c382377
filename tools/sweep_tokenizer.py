"""FPS / kNN-group throughput sweep at the cfg 5 tokenizer shape over the number of co-resident clouds:
python tools/sweep_tokenizer.py [B ...]   (device times: CUDA events, L2 flushed, median of 9)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uniadapter_b200 as ua
from bench import L2Flush
from uniadapter_b200.streams import unit_sphere_clouds

dev = torch.device("cuda:0")
flush = L2Flush(dev)


def med(fn, n=9):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


N, G, k = 1024, 512, 64
for B in [int(a) for a in sys.argv[1:]] or [1, 15, 30, 64, 148, 296, 592, 1184]:
    g = torch.Generator().manual_seed(B)
    xyz = unit_sphere_clouds(B, N, g).to(dev)
    rgb = torch.rand(B, N, 3, generator=g).to(dev)
    _, centers = ua.fps_sample(xyz, G, None)
    f = med(lambda: ua.fps_sample(xyz, G, None))
    f2 = med(lambda: ua.fps_sample(xyz, G, None, pointnet2=True))
    kk = med(lambda: ua.knn_group(xyz, centers, k, rgb))
    k32 = med(lambda: ua.knn_group(xyz, centers, 32))
    by = B * (N * 24 + G * 12 + G * k * 24)
    print(f"B={B:5d}  fps {f:8.1f} us  fps(pn2) {f2:8.1f} us  knn64+rgb {kk:8.1f} us  knn32 {k32:8.1f} us  "
          f"tokenizer {B / (f + kk) * 1e6:10.0f} clouds/s  {by / (f + kk) / 1e3:7.1f} GB/s", flush=True)
