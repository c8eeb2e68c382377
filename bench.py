#!/usr/bin/env python
"""bench.py — adapted point clouds / s of the per-sample test-time hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5]
  (N > 1: launched by torch.distributed.run, one rank per GPU)            prints ONE JSON line on rank 0.

--config names a BASELINE.json configuration (default 2 = configs[1], the one the metric is quoted on):
  1  ULIP-2 PointBERT + DOTA (full covariance), one 1024-point stream, batch 1                 (replicas only at N > 1)
  2  ULIP-2 + MODE-DOTA M=8 + residual text learning, 15 ModelNet40-C streams; weak scaling: 15 streams in lock-step on
     EVERY GPU. The literal "15 streams sharded one per GPU" strong-scaling run rides along as ``scaling_extras``.
  3  OpenShape PPAT, 10 000 coloured points, ball query, MODE-DOTA M=8, K=15                    (replicas only at N > 1)
  4  Uni3D-L, 10 000 points, Objaverse-LVIS cache (K=1156, M=8, D=1024) SHARDED BY CLASS over the N GPUs: strong scaling,
     every rank runs the replicated encoder, the logit exchange runs inside the cache kernel
  5  Uni3D-L geometry at batch 64: tokenizer (FPS + kNN grouping) + head + batched cache step     (replicas only at N > 1)
One step = one sample of every stream of the rank through tokenizer, encoder (PyTorch blocks on the tcgen05 GEMM, timed,
not the target), head, cache predict + fits, (residual learning,) fusion. ``value``: inputs resident in HBM, CUDA events,
L2 flushed before every step, max over ranks. ``e2e``: the same through the engine's public ``step`` with pinned host
clouds in and host logits out. ``roofline``: the north-star kernel of the configuration (its cache step), measured live.
Without --config the line also carries short runs of the other configurations under ``configs`` (``--no-extras`` skips).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "adapted_point_clouds_per_s"
CFG = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}
WORKLOADS = {
    "1": "cfg 1: ULIP-2 PointBERT (random init) + DOTA, synthetic ModelNet40-C stream, 1024 pts, 40 classes, batch 1",
    "2": "ULIP-2 PointBERT (random init) + MODE-DOTA M=8 + res-learning, synthetic ModelNet40-C streams, 1024 pts, 40 classes, batch 1/stream",
    "3": "cfg 3: OpenShape PPAT scaling 4 (random init) + MODE-DOTA M=8, synthetic ScanObjectNN-C streams, 10000 xyz+rgb pts, 15 classes, batch 1/stream",
    "4": "cfg 4: Uni3D-L geometry (random init) + MODE-DOTA M=8, Objaverse-LVIS cache 1156 classes x 1024 dims sharded by class, 10000 pts, batch 1",
    "5": "cfg 5: Uni3D-L tokenizer (512 groups x 64 nn) + head + MODE-DOTA M=8 cache step at batch 64, 1024 xyz+rgb pts, 55 classes (ShapeNet-C)",
}


# ----------------------------------------------------------------------------------------------------------------
# shared measurement plumbing
# ----------------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured copy bandwidth (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """fp32-equivalent peak of the 3xTF32 tensor-core path: measured dense bf16 TFLOP/s (sustained: the GEMMs run
    inside a long step) / 2 for tf32 / 3 products per flop."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p.get("bf16_tflops_sustained", p["bf16_tflops"])) / 6.0, \
            "measured bf16 sustained (MEASURED_PEAKS.json) / 2 (tf32) / 3 (3xTF32): a derived, not a measured, peak"
    return 1400.0 / 6.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md) / 2 / 3"


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed ncu summary of this round
    (profiles/r2_traffic.json: written from ``ncu --set full`` captures, see profiles/README); None when not captured."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        return json.load(open(path)).get(kernel_key, {}).get("dram_bytes")
    except Exception:      # noqa: BLE001
        return None


class L2Flush:
    """Flush of the 126 MB L2 between timed iterations: a 256 MiB memset (evicts everything) followed by a 256 MiB
    read. The read matters: a memset alone leaves the L2 full of DIRTY lines whose write-back to HBM is then charged
    to the timed kernel (profiles/r1_flush_probe.txt)."""

    def __init__(self, dev):
        self.w = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        self.r = torch.ones(64 * 1024 * 1024, dtype=torch.float32, device=dev)
        self.sink = torch.zeros(1, dtype=torch.float32, device=dev)

    def zero_(self):
        self.w.zero_()
        self.sink.copy_(self.r.sum())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:      # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:      # noqa: BLE001
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


class Dist:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max(self, x: float) -> float:
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_steps(D, flush, K, fn):
    """Sum of per-step CUDA-event times (the reference's own event placement, Uni_Adapter.py:379-380,577-579), L2 flushed
    before every step, bracketed by barrier + synchronize, max over ranks. Returns milliseconds for the K steps."""
    evs = []
    D.barrier()
    for i in range(K):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn(i)
        e.record()
        evs.append((s, e))
    D.barrier()
    return D.max(sum(s.elapsed_time(e) for s, e in evs))


def median_us(fn, flush, n=11, warm=3):
    """Device time of one call (CUDA events), L2 flushed before it; a short spin kernel in front of the first event hides
    the host-side launch latency of ``fn`` (the call is queued while the spin still runs)."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


def hbm_roofline(kernel, workload, alg_bytes, us, peak, peak_src, traffic_key=None, **extra):
    ach = alg_bytes / us / 1e3
    r = {"bound": "hbm", "kernel": kernel, "workload": workload, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
         "frac": round(ach / peak, 4), "traffic": ncu_traffic(traffic_key) if traffic_key else None,
         "algorithmic_bytes_per_launch": int(alg_bytes), "launch_us": round(us, 2), "peak_source": peak_src}
    r.update(extra)
    return r


# ----------------------------------------------------------------------------------------------------------------
# kernel-level measurements shared by several configurations
# ----------------------------------------------------------------------------------------------------------------
def cache_step_roofline(dev, flush, S, K, M, D, label):
    """The MODE-DOTA cache pass of a batch-1 sample step (predict + fit + fit on the jittered view) at a given state
    shape: the single-pass kernel (each class tile read and written once: it MOVES 16*S*K*M*D bytes) beside the
    two-launch sequence it replaces (predict+fit, fit: 32*S*K*M*D moved = SURVEY 8d's algorithmic bytes of a sample step)
    and a plain device copy of the state bytes. `achieved` follows the contract (SURVEY 8d's per-unit figure x the units
    one launch processes / the launch's duration); the conservative fraction over the bytes actually moved rides along."""
    from uniadapter_b200.engine import MultiStreamModeDota
    from uniadapter_b200.streams import synthetic_text_features
    peak, src = measured_peaks()
    text = synthetic_text_features(K, D, seed=1).to(dev)
    cache = MultiStreamModeDota(CFG, D, K, text, M, S, dev)
    x = torch.nn.functional.normalize(torch.randn(S, D, device=dev), dim=-1)
    xa = torch.nn.functional.normalize(x + 0.01 * torch.randn(S, D, device=dev), dim=-1)
    g = torch.softmax(100 * x @ text.t(), 1).contiguous()
    out = torch.zeros(S, K, device=dev)
    out3 = out.view(S, 1, K)
    xp = x.half().float().unsqueeze(1)
    single = median_us(lambda: cache.sample_step(x, xa, g, out), flush)

    def two():
        cache.step(xp, x.unsqueeze(1), g.unsqueeze(1), out3)
        cache.step(None, xa.unsqueeze(1), g.unsqueeze(1))
    both = median_us(two, flush)
    one = median_us(lambda: cache.step(xp, x.unsqueeze(1), g.unsqueeze(1), out3), flush)
    by = 16 * S * K * M * D                 # bytes the single pass has to move: every class tile in and out once
    unit = 32 * S * K * M * D               # SURVEY 8d's per-unit figure: one sample step = two fits, state in + out per fit
    src_t = torch.empty(by // 8, device=dev)
    dst_t = torch.empty_like(src_t)
    copy_us = median_us(lambda: dst_t.copy_(src_t), flush)
    r = hbm_roofline("ua_modedota_sample_step_f32 (modedota_sample_kernel): predict + fit + fit, one pass", label, unit,
                     single, peak, src,
                     traffic_key="modedota_sample_kernel_lvis" if K >= 1000 else
                     ("modedota_sample_kernel_cfg2" if (S, K, M, D) == (15, 40, 8, 512) else None),
                     accounting="achieved / frac: SURVEY 8d's algorithmic bytes of a sample step (32*S*K*M*D: predict fused "
                                "into fit #1, two fits, state in + out per fit) over the time of the ONE launch that does the "
                                "whole sample step; the single pass moves half of them (moved_bytes_per_launch), the fraction "
                                "over those bytes is frac_of_moved_bytes",
                     moved_bytes_per_launch=by, achieved_over_moved_bytes=round(by / single / 1e3, 1),
                     frac_of_moved_bytes=round(by / single / 1e3 / peak, 4),
                     two_launch_sequence_us=round(both, 2), two_launch_algorithmic_bytes=2 * by,
                     two_launch_achieved_gbs=round(2 * by / both / 1e3, 1),
                     two_launch_frac=round(2 * by / both / 1e3 / peak, 4),
                     predict_plus_one_fit_us=round(one, 2),
                     predict_plus_one_fit_frac=round(by / one / 1e3 / peak, 4),
                     predict_plus_one_fit_traffic=ncu_traffic("modedota_b1_kernel_lvis") if K >= 1000 else None,
                     plain_copy_same_bytes_us=round(copy_us, 2), frac_of_plain_copy=round(copy_us / single, 4))
    return r


def tokenizer_roofline(dev, flush, B=64, N=1024, G=512, k=64, colored=True, sweep=(64, 148, 592, 1184)):
    """FPS + kNN grouping (the "FPS+kNN GB/s" half of BASELINE.json's metric) at the batch-64 shape of cfg 5 and, because FPS
    is a G-step dependent argmax chain per cloud (latency-bound until several clouds share an SM), over the number of
    clouds in flight: achieved GB/s over SURVEY 8d's algorithmic bytes (cloud in, centres + groups out)."""
    import uniadapter_b200 as ua
    from uniadapter_b200.streams import unit_sphere_clouds
    peak, src = measured_peaks()
    rows = {}
    for b in sorted(set((B,) + tuple(sweep))):
        g = torch.Generator().manual_seed(5 + b)
        xyz = unit_sphere_clouds(b, N, g).to(dev)
        rgb = torch.rand(b, N, 3, generator=g).to(dev) if colored else None
        _, centers = ua.fps_sample(xyz, G, None, pointnet2=colored)
        f = median_us(lambda: ua.fps_sample(xyz, G, None, pointnet2=colored), flush, n=7, warm=2)
        kk = median_us(lambda: ua.knn_group(xyz, centers, k, rgb), flush, n=7, warm=2)
        by = b * (N * (24 if colored else 12) + G * 12 + G * k * (24 if colored else 12))
        rows[b] = {"fps_us": round(f, 1), "knn_group_us": round(kk, 1), "clouds_per_s": round(b / (f + kk) * 1e6),
                   "achieved_gbs": round(by / (f + kk) / 1e3, 1), "frac": round(by / (f + kk) / 1e3 / peak, 4)}
    main = rows[B]
    by = B * (N * (24 if colored else 12) + G * 12 + G * k * (24 if colored else 12))
    return {"workload": f"tokenizer: {B} clouds x {N} {'xyz+rgb' if colored else 'xyz'} points, {G} groups x {k} neighbours",
            "kernels": "ua_fps_f32 (fps_reg_kernel) + ua_knn_group_f32 (knn_group_kernel)", "bound": "hbm",
            "fps_us": main["fps_us"], "knn_group_us": main["knn_group_us"], "algorithmic_bytes": by,
            "achieved": main["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": main["frac"],
            "clouds_per_s": main["clouds_per_s"], "clouds_in_flight_sweep": {str(b): r for b, r in rows.items()},
            "peak_source": src,
            "note": "FPS is latency-bound (serial argmax chain per cloud), kNN selection is issue-bound: the HBM fraction is "
                    "reported, clouds/s is the meaningful figure (DESIGN.md)"}


# ----------------------------------------------------------------------------------------------------------------
# configurations
# ----------------------------------------------------------------------------------------------------------------
class Cfg2:
    """ULIP-2 + MODE-DOTA M=8 + residual learning, S streams in lock-step per GPU (weak scaling over GPUs)."""
    key, N, K, D, M = "2", 1024, 40, 512, 8
    scaling = "weak"

    def __init__(self, D_, args):
        from uniadapter_b200.encoders import build_encoder
        from uniadapter_b200.engine import StreamEngine
        from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
        self.S = args.streams
        dev = D_.dev
        self.encoder = build_encoder('ulip', seed=0, device=dev)
        self.text = synthetic_text_features(self.K, self.D, seed=0)
        self.engine = StreamEngine(self.encoder, 'ulip', self.text, self.S, self.N, CFG, mode_M=self.M, res_learning=True,
                                   device=dev, use_graph=not args.no_graph, seed=42,
                                   stream_ids=[D_.rank * self.S + s for s in range(self.S)])
        g = torch.Generator().manual_seed(42 + D_.rank)
        self.host = [unit_sphere_clouds(self.S, self.N, g).pin_memory() for _ in range(8)]
        self.resident = [h.to(dev) for h in self.host]
        self.clouds_per_step = self.S * D_.world
        self.h2d, self.d2h = self.S * self.N * 12, self.S * self.K * 4
        self.parallelism = f"stream-per-GPU x{D_.world}: {self.S} streams in lock-step on every GPU (no collective)"

    def step_device(self, i):
        self.engine.step_device(self.resident[i % 8])

    def step_e2e(self, i):
        self.engine.step(self.host[i % 8])

    def launches(self):
        return self.engine.launches_per_step()

    def roofline(self, dev, flush):
        r = cache_step_roofline(dev, flush, self.S, self.K, self.M, self.D,
                                f"cfg 2 cache step: {self.S} streams x K=40 M=8 D=512 (state 2 x {self.S * self.K * self.M * self.D * 4 / 1e6:.1f} MB, L2-resident between steps in production; flushed here)")
        r["step_kernels"] = step_kernel_table(self.engine, self.resident, flush, self)
        return r


def step_kernel_table(engine, resident, flush, cfg):
    """Per-kernel device times of one eager step (events around every library call), with the encoder GEMM's tensor
    figures for context (the transformer blocks are timed, not the target)."""
    from uniadapter_b200 import _lib
    was = engine.use_graph
    engine.use_graph = False
    tl = _lib.enable_kernel_timing(True)
    reps = 3
    try:
        for i in range(reps):
            flush.zero_()
            engine.step_device(resident[i % len(resident)])
        torch.cuda.synchronize()
        summ = tl.summary()
    finally:
        _lib.enable_kernel_timing(False)
        engine.use_graph = was
    tensor_peak, tensor_src = measured_tensor_peak()
    kern = {}
    for n, (c, tot, m) in summ.items():
        row = {"calls_per_step": c // reps, "mean_us": round(m, 2), "total_us_per_step": round(tot / reps, 1)}
        if n in tl.flops:
            ach = tl.flops[n] / (tot * 1e-6) / 1e12
            row.update(achieved_tflops_fp32_equiv=round(ach, 2), tensor_frac=round(ach / tensor_peak, 4),
                       tensor_peak=round(tensor_peak, 1), tensor_peak_source=tensor_src,
                       traffic_largest_launch=ncu_traffic("gemm_tf32x3_conv3"))
        kern[n] = row
    return kern


class Cfg1:
    """ULIP-2 + DOTA (full covariance, K=40, D=512): one stream, the step as one CUDA-graph replay (engine.DotaEngine)."""
    key, N, K, D = "1", 1024, 40, 512
    scaling = "weak"

    def __init__(self, D_, args):
        from uniadapter_b200.encoders import build_encoder
        from uniadapter_b200.engine import DotaEngine
        from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
        dev = D_.dev
        self.encoder = build_encoder('ulip', seed=0, device=dev)
        self.text = synthetic_text_features(self.K, self.D, seed=0)
        self.engine = DotaEngine(self.encoder, 'ulip', self.text, self.N, CFG, device=dev, use_graph=not args.no_graph,
                                 seed=42, stream_id=D_.rank)
        g = torch.Generator().manual_seed(42 + D_.rank)
        self.host = [unit_sphere_clouds(1, self.N, g).pin_memory() for _ in range(8)]
        self.resident = [h.to(dev) for h in self.host]
        self.clouds_per_step = D_.world
        self.h2d, self.d2h = self.N * 12, self.K * 4
        self.parallelism = f"replicas only x{D_.world}: one independent stream per GPU (DOTA does not shard)"

    def step_device(self, i):
        self.engine.step_device(self.resident[i % 8])

    def step_e2e(self, i):
        self.engine.step(self.host[i % 8])

    def launches(self):
        from uniadapter_b200 import _lib
        was = self.engine.use_graph
        self.engine.use_graph = False
        _lib.reset_launch_count()
        self.engine.step(self.host[0])
        torch.cuda.synchronize()
        n = _lib.launch_count()
        self.engine.use_graph = was
        return n

    def roofline(self, dev, flush):
        import uniadapter_b200 as ua
        peak, src = measured_peaks()
        K, D = self.K, self.D
        a = ua.DOTA(CFG, D, K, torch.full((D, K), 0.001), device=dev)
        x = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
        y = torch.softmax(torch.randn(1, K, device=dev), 1)
        fit = median_us(lambda: a.fit(x, y), flush)
        upd = median_us(lambda: a.update(), flush)
        prd = median_us(lambda: a.predict(x.half()), flush)
        by = 8 * K * D * D + 4 * D * D
        return hbm_roofline("ua_dota_fit_f32 (dota_sigma_kernel + dota_mean_kernel)",
                            "cfg 1 DOTA fit: K=40 D=512, Sigma [K,D,D] read + written once, class mean fused", by, fit, peak, src,
                            traffic_key="dota_sigma_kernel_cfg1", update_inverse_us=round(upd, 2), predict_f16_us=round(prd, 2))


class Cfg3:
    """OpenShape PPAT + MODE-DOTA M=8 (K=15, D=1280), 10 000 coloured points, S streams in lock-step."""
    key, N, K, D, M = "3", 10000, 15, 1280, 8
    scaling = "weak"

    def __init__(self, D_, args):
        from uniadapter_b200.encoders import build_encoder
        from uniadapter_b200.engine import StreamEngine
        from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
        dev = D_.dev
        self.S = 4
        self.encoder = build_encoder('openshape', seed=0, device=dev)
        self.text = synthetic_text_features(self.K, self.D, seed=0)
        self.engine = StreamEngine(self.encoder, 'openshape', self.text, self.S, self.N, CFG, mode_M=self.M,
                                   res_learning=False, device=dev, use_graph=not args.no_graph, seed=42,
                                   stream_ids=[D_.rank * self.S + s for s in range(self.S)])
        g = torch.Generator().manual_seed(42 + D_.rank)
        self.host = [unit_sphere_clouds(self.S, self.N, g).pin_memory() for _ in range(4)]
        self.rgb = [torch.rand(self.S, self.N, 3, generator=g).pin_memory() for _ in range(4)]
        self.resident = [h.to(dev) for h in self.host]
        self.engine.rgb.copy_(self.rgb[0])
        self.clouds_per_step = self.S * D_.world
        self.h2d, self.d2h = self.S * self.N * 24, self.S * self.K * 4
        self.parallelism = f"replicas only x{D_.world}: {self.S} streams in lock-step per GPU"

    def step_device(self, i):
        self.engine.step_device(self.resident[i % 4])

    def step_e2e(self, i):
        self.engine.step(self.host[i % 4], self.rgb[i % 4])

    def launches(self):
        return self.engine.launches_per_step()

    def roofline(self, dev, flush):
        import uniadapter_b200 as ua
        from uniadapter_b200.streams import unit_sphere_clouds
        r = cache_step_roofline(dev, flush, self.S, self.K, self.M, self.D, f"cfg 3 cache step: {self.S} streams x K=15 M=8 D=1280")
        peak, src = measured_peaks()
        B, N, S_, ns = 2 * self.S, self.N, 384, 64
        g = torch.Generator().manual_seed(9)
        xyz = unit_sphere_clouds(B, N, g).to(dev)
        pts = torch.cat((xyz, torch.rand(B, N, 3, generator=g).to(dev)), -1).contiguous()
        st = torch.randint(0, N, (B,), generator=g).to(dev)
        _, cen = ua.fps_sample(xyz, S_, st)
        f = median_us(lambda: ua.fps_sample(xyz, S_, st), flush, n=7, warm=2)
        b = median_us(lambda: ua.ball_group(xyz, cen, 0.2, ns, pts), flush, n=7, warm=2)
        by = B * (N * 24 + S_ * 12 + S_ * ns * 36)
        r["tokenizer"] = {"workload": f"FPS + ball query (r=0.2, 64 samples) on {B} clouds x 10000 xyz+rgb points, 384 patches",
                          "kernels": "ua_fps_f32 (fps_cluster_kernel) + ua_ball_group_f32", "fps_us": round(f, 1),
                          "ball_group_us": round(b, 1), "algorithmic_bytes": by, "achieved": round(by / (f + b) / 1e3, 1),
                          "unit": "GB/s", "peak": peak, "frac": round(by / (f + b) / 1e3 / peak, 4), "peak_source": src}
        return r


class Cfg4:
    """Uni3D-L + Objaverse-LVIS MODE-DOTA cache sharded by class over the GPUs of the job (strong scaling)."""
    key, N, K, D, M = "4", 10000, 1156, 1024, 8
    scaling = "strong"

    def __init__(self, D_, args):
        from uniadapter_b200.encoders import build_encoder
        from uniadapter_b200.engine import ShardedSampleEngine
        from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
        dev = D_.dev
        self.encoder = build_encoder('uni3d', seed=0, device=dev)
        self.text = synthetic_text_features(self.K, self.D, seed=0)
        self.engine = ShardedSampleEngine(self.encoder, 'uni3d', self.text, self.N, CFG, mode_M=self.M, device=dev,
                                          use_graph=not args.no_graph, seed=42, stream_id=0)
        g = torch.Generator().manual_seed(42)        # the SAME stream on every rank (replicated encoder, sharded cache)
        self.host = [unit_sphere_clouds(1, self.N, g).pin_memory() for _ in range(8)]
        self.resident = [h.to(dev) for h in self.host]
        self.clouds_per_step = 1
        self.h2d, self.d2h = self.N * 12, self.K * 4
        self.world = D_.world
        self.parallelism = (f"class-sharded cache x{D_.world}: {-(-self.K // D_.world)} classes per GPU, replicated encoder, logit "
                            "exchange over NVLink peer memory inside the cache kernel (2 flag exchanges per step)")

    def step_device(self, i):
        self.engine.step_device(self.resident[i % 8])

    def step_e2e(self, i):
        self.engine.step(self.host[i % 8])

    def launches(self):
        from uniadapter_b200 import _lib
        was = self.engine.use_graph
        self.engine.use_graph = False
        _lib.reset_launch_count()
        self.engine.step_device(self.resident[0])
        torch.cuda.synchronize()
        n = _lib.launch_count()
        self.engine.use_graph = was
        return n

    def roofline(self, dev, flush):
        r = cache_step_roofline(dev, flush, 1, self.K, self.M, self.D,
                                "cfg 4 cache step, unsharded: K=1156 M=8 D=1024 (151.5 MB in + out per sample step, > L2)")
        r["sharded_step"] = sharded_step_probe(self, dev, flush)
        return r


def sharded_step_probe(cfg, dev, flush):
    """Device time of the adapter part of the cfg 4 step at this world size (head on the local text rows + the fused
    class-sharded kernel), L2 flushed first, beside the single-GPU adapter step (head + single-pass cache step + fusion)."""
    import torch.distributed as dist
    import uniadapter_b200 as ua
    fused = cfg.engine.sharded
    K, D, M = cfg.K, cfg.D, cfg.M
    emb = torch.randn(2, D, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(emb, 0)

    def sync_all():
        torch.cuda.synchronize()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()

    g = torch.cuda.CUDAGraph()
    fused.enqueue(emb)
    sync_all()
    with torch.cuda.graph(g):
        fused.enqueue(emb)

    def chain(n, with_step):
        """n x [L2 flush, step] queued back to back (no host synchronisation in between: the ranks pace one another
        through the exchange itself, as they do in the real loop), device time of the whole chain."""
        sync_all()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2000000)
        s.record()
        for _ in range(n):
            flush.zero_()
            if with_step:
                g.replay()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) * 1e3

    n = 20
    chain(3, True)
    per_step = sorted((chain(n, True) - chain(n, False)) / n for _ in range(3))[1]
    fused.check()
    t = torch.tensor([per_step], device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {"world": cfg.world, "classes_per_rank": fused.mine.Kp, "fused_sharded_adapter_step_us": round(float(t), 2),
           "what": "ua_head_f32 on the local text rows + ua_modedota_sharded_step_f32 (exchange, predict + 2 fits, exchange, "
                   "fusion) as one CUDA-graph replay; 20 x [L2 flush, step] queued back to back minus 20 x [L2 flush], per "
                   "step, median of 3, max over ranks (steady state: no host-side start skew between the ranks)"}
    # single-GPU adapter step for comparison (every rank measures it locally; rank 0's figure is reported)
    text = cfg.text.to(dev)
    full = ua.DOTA_mix(CFG, D, K, text.t().contiguous(), num_modes=M, device=dev)
    x0, x1 = emb[0:1].contiguous(), emb[1:2].contiguous()

    def single():
        feats, clip_logits, _, prob, _ = ua.zero_shot_head(x0, text)
        feats_aug = ua.zero_shot_head(x1, text)[0]
        dl = full.sample_step(feats, feats_aug, prob)
        ua.fuse_logits(clip_logits, dl, full.c, CFG['rho'], CFG['eta'], 1, 'mode_dota')
    single()
    torch.cuda.synchronize()
    g1 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g1):
        single()
    out["single_gpu_adapter_step_us"] = round(median_us(lambda: g1.replay(), flush, n=15), 2)
    return out


class Cfg5:
    """Uni3D-L tokenizer + head + batched MODE-DOTA cache step at batch 64 (K=55, D=1024): throughput of the hot path
    without the transformer blocks (BASELINE cfg 5 names tokenizer + cache step)."""
    key, N, K, D, M, B, G, k = "5", 1024, 55, 1024, 8, 64, 512, 64
    scaling = "weak"

    def __init__(self, D_, args):
        import uniadapter_b200 as ua
        from uniadapter_b200.head import HeadPlan
        from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
        dev = D_.dev
        self.dev = dev
        g = torch.Generator().manual_seed(42 + D_.rank)
        self.host = [(unit_sphere_clouds(self.B, self.N, g).pin_memory(), torch.rand(self.B, self.N, 3, generator=g).pin_memory())
                     for _ in range(4)]
        self.resident = [(a.to(dev), b.to(dev)) for a, b in self.host]
        self.xyz, self.rgb = torch.zeros(self.B, self.N, 3, device=dev), torch.zeros(self.B, self.N, 3, device=dev)
        torch.manual_seed(0)
        # cfg 5 names tokenizer + cache step: the encoder between them (mini-PointNet + blocks) is NOT in this step; a fixed
        # projection of the sampled centres stands in for it so that the head and the cache see data-dependent features
        self.proj = torch.randn(3, self.D, device=dev)
        self.text = synthetic_text_features(self.K, self.D, seed=0).to(dev)
        self.head = HeadPlan(self.text)
        self.model = ua.DOTA_mix(CFG, self.D, self.K, self.text.t().contiguous(), num_modes=self.M, device=dev)
        self.final = torch.zeros(self.B, self.K, device=dev)
        self._host_out = torch.empty(self.B, self.K).pin_memory()
        self.clouds_per_step = self.B * D_.world
        self.h2d, self.d2h = self.B * self.N * 24, self.B * self.K * 4
        self.parallelism = f"replicas only x{D_.world}: one batch-64 stream per GPU"
        self.graph = None
        self.use_graph = not args.no_graph
        self.i = 0

    @torch.no_grad()
    def _body(self):
        import uniadapter_b200 as ua
        _, centers = ua.fps_sample(self.xyz, self.G, None, want_idx=False, pointnet2=True)
        _, _, self.feat = ua.knn_group(self.xyz, centers, self.k, self.rgb, want_neigh=False)      # (B,G,k,6) written to HBM
        emb = centers.amax(1) @ self.proj                                 # encoder stand-in (not part of cfg 5's scope)
        feats, logits, _, prob, _ = self.head(emb)
        dl = self.model.predict_then_fit(feats.half().float(), feats, prob)
        final, _, _ = ua.fuse_logits(logits, dl, self.model.c, CFG['rho'], CFG['eta'], self.B, 'mode_dota')
        self.final.copy_(final)

    def _run(self):
        if self.use_graph and self.i >= 2:
            if self.graph is None:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._body()
            self.graph.replay()
        else:
            self._body()
        self.i += 1

    def step_device(self, i):
        a, b = self.resident[i % 4]
        self.xyz.copy_(a)
        self.rgb.copy_(b)
        self._run()

    def step_e2e(self, i):
        a, b = self.host[i % 4]
        self.xyz.copy_(a, non_blocking=True)
        self.rgb.copy_(b, non_blocking=True)
        self._run()
        self._host_out.copy_(self.final, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def launches(self):
        from uniadapter_b200 import _lib
        was, self.use_graph = self.use_graph, False
        _lib.reset_launch_count()
        self.step_device(0)
        torch.cuda.synchronize()
        n = _lib.launch_count()
        self.use_graph = was
        return n

    def roofline(self, dev, flush):
        peak, src = measured_peaks()
        r = tokenizer_roofline(dev, flush)
        B, K, M, D = self.B, self.K, self.M, self.D
        x = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=-1)
        gm = torch.softmax(100 * x @ self.text.t(), 1)
        us = median_us(lambda: self.model.predict_then_fit(x, x, gm), flush)
        by = 16 * K * M * D + 8 * B * D
        r["cache_step_b64"] = hbm_roofline("ua_modedota_step_f32 (modedota_batch_kernel): predict + fit at batch 64",
                                           "K=55 M=8 D=1024 B=64", by, us, peak, src,
                                           note="B*K*M*D = 28.8 M (row, mode, d) terms against 14.4 MB of state: issue-bound, not HBM-bound")
        hp = median_us(lambda: self.head(x), flush)
        r["head_b64_tcgen05"] = {"kernel": "ua_head_prepare_f32 + ua_gemm_tf32x3_f32 + ua_row_stats_f32 (HeadPlan)", "us": round(hp, 2),
                                 "flops": 2 * B * K * D}
        return r


CONFIGS = {"1": Cfg1, "2": Cfg2, "3": Cfg3, "4": Cfg4, "5": Cfg5}


# ----------------------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the reference's OWN CPU implementation on the host cores
# ----------------------------------------------------------------------------------------------------------------
def reference_available():
    from oracle import reference_loader as R
    return R.available()


def cpu_reference_rate(key, steps, warmup, threads):
    """Clouds/s of the reference's CPU path on this host, one stream, batch 1, sequential samples (the reference adapts
    one sample at a time). cfg 1-3: the reference's own modules (oracle/ref_loop.py: PointTransformer / PPAT,
    get_logits_wrapper, DOTA / DOTA_mix, compute_text_alignment_loss, Adam) imported from oracle/_ref -> kind "reference".
    cfg 4/5: Uni3D needs the un-vendored pointnet2_ops and timm, so the oracle port runs instead -> kind "port"."""
    from oracle import synth
    torch.set_num_threads(threads)
    if key in ("1", "2", "3") and reference_available():
        from oracle import ref_loop
        if key == "3":
            model = ref_loop.openshape_reference_model(12, 384, 0)
            text = torch.from_numpy(synth.unit_rows(15, 1280, 1))
            stream = ref_loop.ReferenceStream(model, 'openshape', text, CFG, 1280, 8, False)
            n = 10000
        else:
            model, _, _ = ref_loop.ulip_reference_model(12, 0)
            text = torch.from_numpy(synth.unit_rows(40, 512, 1))
            stream = ref_loop.ReferenceStream(model, 'ulip', text, CFG, 512, 8 if key == "2" else 0, key == "2")
            n = 1024
        pcs = torch.from_numpy(synth.cloud(steps + warmup, n, 5))
        rgb = torch.rand(1, n, 3) if key == "3" else torch.ones(1, n, 3)
        for i in range(warmup):
            stream.step(pcs[i:i + 1], rgb)
        t0 = time.perf_counter()
        for i in range(warmup, warmup + steps):
            stream.step(pcs[i:i + 1], rgb)
        dt = time.perf_counter() - t0
        return steps / dt, dt / steps, "reference", 1
    # oracle port (C tokenizer + numpy adapters + this repo's module definitions on the CPU)
    from oracle.cpu_pipeline import CpuStream, cpu_encoder_like
    if key == "5":
        from oracle import adapters as A
        from oracle import tokenizer as T
        B, N, G, k, K, D, M = 64, 1024, 512, 64, 55, 1024, 8
        text = synth.unit_rows(K, D, 1)
        model = A.ModeDota(CFG, D, K, text.T, M)
        xyz, rgb = synth.cloud(B * (steps + warmup), N, 5).reshape(steps + warmup, B, N, 3), synth.uniform((B, N, 3), 6)
        x, _, _ = synth.features(steps + warmup, B, D, text, 7)

        def one(i):
            T.group_knn(xyz[i], G, k, rgb=rgb, threads=threads, pointnet2=True)
            h = A.head(x[i], text)
            model.predict(h["xnorm"].astype("float16").astype("float32"))
            model.fit(h["xnorm"], h["prob"])
        for i in range(warmup):
            one(i)
        t0 = time.perf_counter()
        for i in range(warmup, warmup + steps):
            one(i)
        dt = time.perf_counter() - t0
        return B * steps / dt, dt / steps, "port", B
    from uniadapter_b200.encoders import Uni3DEncoder, UlipPointBert, OpenShapePPAT
    torch.manual_seed(0)
    fam, enc_m, K, D, n, M, res = {"1": ('ulip', UlipPointBert, 40, 512, 1024, 0, False), "2": ('ulip', UlipPointBert, 40, 512, 1024, 8, True),
                                   "3": ('openshape', OpenShapePPAT, 15, 1280, 10000, 8, False),
                                   "4": ('uni3d', Uni3DEncoder, 1156, 1024, 10000, 8, False)}[key]
    enc = cpu_encoder_like(enc_m().eval(), threads=threads)
    text = synth.unit_rows(K, D, 1)
    stream = CpuStream(enc, fam, text, CFG, 'mode_dota' if M else 'dota', max(M, 1), res)
    pcs = torch.from_numpy(synth.cloud(steps + warmup, n, 5))
    rgb = torch.ones(1, n, 3)
    for i in range(warmup):
        stream.step(pcs[i:i + 1], rgb)
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        stream.step(pcs[i:i + 1], rgb)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps, "port", 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    key = args.config or "2"
    # bounded sample: the same number of warm-up steps as the repo arm where that stays within minutes
    steps = max(1, min(args.steps, 30 if key in ("1", "2", "5") else 6))
    warm = max(1, min(max(args.warmup, 3), 5 if key in ("1", "2", "5") else 2))
    rate, sec, kind, clouds = cpu_reference_rate(key, steps, warm, threads)
    sample = (f"{steps} sequential {'batches of 64' if clouds > 1 else 'samples'} of one stream after {warm} warm-up, all {threads} host "
              f"threads; " + ("the reference's own modules from oracle/_ref (oracle/ref_loop.py)" if kind == "reference" else
                              "oracle port (the reference needs un-vendored pointnet2_ops / timm here)"))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "clouds/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": CONFIGS[key].scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[key], "streams_per_gpu": 1,
                       "note": "the reference adapts one sample of one stream at a time; its encoder depth, shapes and adapter are the repo arm's"},
            "cpu_baseline": {"value": rate, "unit": "clouds/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# the repo arm
# ----------------------------------------------------------------------------------------------------------------
def measure_config(D, args, key, flush, steps, warmup, with_roofline=True, with_kernel_table=True):
    cfg = CONFIGS[key](D, args)
    launches = cfg.launches()                    # eager step 0
    for i in range(max(warmup, 3)):
        cfg.step_e2e(i)
    launches = cfg.launches()                    # steady-state step, counted by the library itself
    torch.cuda.synchronize()
    ms_res = timed_steps(D, flush, steps, cfg.step_device)
    ms_e2e = timed_steps(D, flush, steps, cfg.step_e2e)
    out = {"value": round(cfg.clouds_per_step * steps / (ms_res * 1e-3), 2), "ms_per_step": round(ms_res / steps, 4),
           "e2e": {"value": round(cfg.clouds_per_step * steps / (ms_e2e * 1e-3), 2), "unit": "clouds/s",
                   "h2d_bytes_per_step": cfg.h2d, "d2h_bytes_per_step": cfg.d2h, "ms_per_step": round(ms_e2e / steps, 4)},
           "gpu_launches": launches * steps * 2, "gpu_launches_per_step": launches,
           "clouds_per_step": cfg.clouds_per_step, "parallelism": cfg.parallelism, "scaling": cfg.scaling}
    if key == "2":          # single-stream latency beside the lock-step throughput
        out["streams_per_gpu"] = cfg.S
    if with_roofline:
        try:
            out["roofline"] = cfg.roofline(D.dev, flush)
        except Exception as exc:      # noqa: BLE001  (auxiliary: must not cost the headline line)
            out["roofline"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if hasattr(cfg, "engine") and hasattr(cfg.engine, "check"):
        cfg.engine.check()
    return out, cfg


def cfg2_strong_scaling(D, args, flush, steps):
    """BASELINE configs[1] as written: 15 corruption streams sharded over the GPUs (stream s on rank s mod P: 8 + 7 at
    P = 8), every rank advancing its own streams in lock-step; the job finishes when the slowest rank does."""
    from uniadapter_b200 import parallel
    from uniadapter_b200.encoders import build_encoder
    from uniadapter_b200.engine import StreamEngine
    from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds
    mine = parallel.assign_streams(15, D.world, D.rank)
    dev = D.dev
    enc = build_encoder('ulip', seed=0, device=dev)
    eng = StreamEngine(enc, 'ulip', synthetic_text_features(40, 512, seed=0), len(mine), 1024, CFG, mode_M=8,
                       res_learning=True, device=dev, use_graph=not args.no_graph, seed=42, stream_ids=mine)
    g = torch.Generator().manual_seed(7 + D.rank)
    res = [unit_sphere_clouds(len(mine), 1024, g).to(dev) for _ in range(4)]
    for i in range(4):
        eng.step_device(res[i % 4])
    ms = timed_steps(D, flush, steps, lambda i: eng.step_device(res[i % 4]))
    return {"what": "BASELINE configs[1] literally: 15 streams split over the GPUs (stream s -> rank s mod P), strong scaling",
            "streams_per_rank_max": -(-15 // D.world), "value": round(15 * steps / (ms * 1e-3), 2), "unit": "clouds/s",
            "ms_per_step": round(ms / steps, 4), "ideal_speedup_at_8": 7.5}


def run_ours(args):
    D = Dist()
    flush = L2Flush(D.dev)
    key = args.config or "2"
    K, W = args.steps, max(args.warmup, 3)
    with ClockSampler(D.local) as clocks:
        main, cfg = measure_config(D, args, key, flush, K, W)
    line = {"metric": METRIC, "value": main["value"], "unit": "clouds/s", "n_gpus": D.world, "steps": K, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": main["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[key], "baseline_config": int(key), "clouds_per_step": main["clouds_per_step"],
                       "l2": "flushed before every timed step (256 MiB memset, then 256 MiB read so that no dirty lines remain)",
                       "cuda_graph": not args.no_graph, "parallelism": main["parallelism"]},
            "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "gpu_launches_per_step": main["gpu_launches_per_step"],
            "roofline": main.get("roofline"), "clocks": clocks.summary()}
    if key == "2":
        line["config"]["streams_per_gpu"] = cfg.S
    del cfg
    torch.cuda.empty_cache()
    extras = not args.no_extras and args.config is None
    if extras:
        # ---- single-stream latency of the headline configuration (the reference adapts one stream at a time) ----------
        try:
            a1 = argparse.Namespace(**vars(args))
            a1.streams = 1
            one, c1 = measure_config(D, a1, "2", flush, max(4, K // 2), 3, with_roofline=False)
            line["single_stream"] = {"value": one["value"], "ms_per_step": one["ms_per_step"], "e2e": one["e2e"]["value"],
                                     "what": "the same workload with ONE stream per GPU: per-sample latency of the graph-captured step"}
            del c1
        except Exception as exc:      # noqa: BLE001
            line["single_stream"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        torch.cuda.empty_cache()
        # ---- the other BASELINE configurations, short runs (same code path as --config N) ------------------------------
        line["configs"] = {}
        for k2 in ("1", "3", "4", "5"):
            try:
                res, c2 = measure_config(D, args, k2, flush, max(4, min(K, 10)), 3)
                res["workload"] = WORKLOADS[k2]
                line["configs"][k2] = res
                del c2
            except Exception as exc:      # noqa: BLE001
                line["configs"][k2] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()
        if D.world > 1:
            try:
                line["scaling_extras"] = {"cfg2_strong": cfg2_strong_scaling(D, args, flush, max(4, min(K, 10)))}
            except Exception as exc:      # noqa: BLE001
                line["scaling_extras"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if D.rank == 0:
        if D.world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            try:
                rate, sec, kind, _ = cpu_reference_rate(key, 4 if key in ("1", "2", "5") else 2, 1, threads)
                line["cpu_baseline"] = {"value": round(rate, 4), "unit": "clouds/s", "cores": threads, "kind": kind,
                                        "sample": f"{4 if key in ('1', '2', '5') else 2} sequential samples of one stream after 1 warm-up, same workload "
                                                  + ("(the reference's own modules, oracle/ref_loop.py)" if kind == "reference" else "(oracle port)")}
            except Exception as exc:      # noqa: BLE001
                line["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        print(json.dumps(line))
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=list(CONFIGS), default=None, help="BASELINE.json configuration (default: 2, plus "
                    "short runs of the others)")
    ap.add_argument("--streams", type=int, default=15, help="corruption streams advanced in lock-step per GPU (cfg 2)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""      # the CPU arm: the reference places its state on cuda when it sees one
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
