#!/usr/bin/env python
"""bench.py — adapted point clouds / s of the per-sample test-time hot path on B200.

Workload (BASELINE.json configs[1]): ULIP-2 PointBERT (random init) + MODE-DOTA (M=8) + residual text learning on
synthetic ModelNet40-C-shaped streams (1024 points, 40 classes, batch 1 per stream); S independent corruption streams
advance in lock-step on every GPU (one process per GPU, no data-path collective: "weak" scaling).
One step = one sample of every stream on the rank = S clouds: tokenizer x2, encoder x2 (PyTorch, timed, not the
target), head, cache predict + 2 fits, 11 alignment-loss evaluations / 10 Adam steps, fusion.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU)

Prints ONE JSON line on rank 0.
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
N_POINTS, N_CLASSES, FEAT_DIM, MODES = 1024, 40, 512, 8
# dram__bytes_read.sum + dram__bytes_write.sum of one LVIS-scale launch (ncu --set full, profiles/r1_modedota_lvis.txt):
# 75.9 MB read + 17.2 MB written inside the kernel; the rest of the 75.8 MB of output is still dirty in L2 at exit.
LVIS_DRAM_TRAFFIC_BYTES = 93_060_000
GEMM_DRAM_TRAFFIC_BYTES = 1_468_000_000   # largest GEMM launch of the step (group-encoder conv3: 245 760 x 512 x 256, split output): 520.7 MB read + 947.3 MB written (ncu --set full, profiles/r1_ncu_hot_kernels.txt)
WORKLOAD = "ULIP-2 PointBERT (random init) + MODE-DOTA M=8 + res-learning, synthetic ModelNet40-C streams, 1024 pts, 40 classes, batch 1/stream"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """fp32-equivalent peak of the 3xTF32 tensor-core path: measured dense bf16 TFLOP/s (sustained: the GEMMs run
    inside a long step) / 2 for tf32 / 3 products per flop."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p.get("bf16_tflops_sustained", p["bf16_tflops"])) / 6.0, \
            "measured bf16 sustained (MEASURED_PEAKS.json) / 2 (tf32) / 3 (3xTF32)"
    return 1400.0 / 6.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md) / 2 / 3"


class L2Flush:
    """Flush of the 126 MB L2 between timed iterations: a 256 MiB memset (evicts everything) followed by a 256 MiB
    read. The read matters: a memset alone leaves the L2 full of DIRTY lines whose write-back to HBM is then charged
    to the timed kernel (a plain 152 MB device copy runs at 4.9 TB/s after a memset-only flush and at 6.2 TB/s after
    memset + read; profiles/r1_flush_probe.txt)."""

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
        except Exception:
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
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_rate(steps, warmup, threads):
    """Clouds/s of the reference's CPU path (oracle port) on this host: one stream, batch 1, sequential samples."""
    from oracle import synth
    from oracle.cpu_pipeline import CpuStream, cpu_encoder_like
    from uniadapter_b200.encoders import UlipPointBert
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    enc = cpu_encoder_like(UlipPointBert().eval(), threads=threads)
    text = synth.unit_rows(N_CLASSES, FEAT_DIM, 1)
    stream = CpuStream(enc, 'ulip', text, CFG, 'mode_dota', MODES, True)
    pcs = torch.from_numpy(synth.cloud(steps + warmup, N_POINTS, 5))
    rgb = torch.ones(1, N_POINTS, 3)
    for i in range(warmup):
        stream.step(pcs[i:i + 1], rgb)
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        stream.step(pcs[i:i + 1], rgb)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 30))
    warm = max(1, min(args.warmup, 3))
    rate, sec = cpu_reference_rate(steps, warm, threads)
    sample = f"{steps} sequential samples of one stream after {warm} warm-up (the reference adapts one sample at a time)"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "clouds/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "streams_per_gpu": 1},
            "cpu_baseline": {"value": rate, "unit": "clouds/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    import torch.distributed as dist
    import uniadapter_b200 as ua
    from uniadapter_b200 import _lib
    from uniadapter_b200.encoders import build_encoder
    from uniadapter_b200.engine import StreamEngine
    from uniadapter_b200.streams import synthetic_text_features, unit_sphere_clouds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, K, W = args.streams, args.steps, max(args.warmup, 3)

    encoder = build_encoder('ulip', seed=0, device=dev)
    text = synthetic_text_features(N_CLASSES, FEAT_DIM, seed=0)
    engine = StreamEngine(encoder, 'ulip', text, S, N_POINTS, CFG, mode_M=MODES, res_learning=True, device=dev,
                          use_graph=not args.no_graph, seed=42 + rank)
    g = torch.Generator().manual_seed(42 + rank)
    pool = 8
    host = [unit_sphere_clouds(S, N_POINTS, g).pin_memory() for _ in range(pool)]
    resident = [h.to(dev) for h in host]
    flush = L2Flush(dev)

    launches = engine.launches_per_step()      # eager step 0 (no residual learning yet)
    for i in range(W):
        engine.step(host[i % pool])
    launches = engine.launches_per_step()      # steady-state step (counted by the library)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        """Sum of per-step CUDA-event times (the reference's own event placement), L2 flushed before every step."""
        evs = []
        barrier()
        for i in range(K):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(i)
            e.record()
            evs.append((s, e))
        barrier()
        total_ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with ClockSampler(local) as clocks:
        ms_resident = timed(lambda i: engine.step_device(resident[i % pool]))
        ms_e2e = timed(lambda i: engine.step(host[i % pool]))
    value = world * S * K / (ms_resident * 1e-3)
    e2e = world * S * K / (ms_e2e * 1e-3)

    # ---- per-kernel times of one eager step (events around every library call) -> dominant kernel, roofline ----
    peak, peak_src = measured_peaks()
    engine.use_graph = False
    tl = _lib.enable_kernel_timing(True)
    reps = 5
    for i in range(reps):
        flush.zero_()
        engine.step_device(resident[i % pool])
    torch.cuda.synchronize()
    summ = tl.summary()
    _lib.enable_kernel_timing(False)
    G, kk = 512, 32
    V = 2 if engine.batch_views else 1      # clouds per stream in one tokenizer / encoder launch (sample + jittered view)
    alg_bytes = {   # algorithmic bytes per launch
        "ua_fps_f32": V * S * (N_POINTS * 12 + G * 12),
        "ua_knn_group_f32": V * S * (N_POINTS * 12 + G * 12 + G * kk * 12),
        "ua_head_f32": 4 * S * (FEAT_DIM + FEAT_DIM * N_CLASSES + N_CLASSES),
        "ua_modedota_step_f32": 16 * S * N_CLASSES * MODES * FEAT_DIM,
        "ua_fuse_logits_f32": 4 * S * 3 * N_CLASSES,
        # one call = 10 Adam steps = 42 launches; forward and backward each read mu and var once (L2-resident)
        "ua_residual_learn_f32": 10 * 2 * 8 * S * N_CLASSES * MODES * FEAT_DIM,
    }
    notes = {
        "ua_fps_f32": "FPS is a G-step serial argmax chain per cloud: latency-bound by construction (DESIGN.md)",
        "ua_residual_learn_f32": "one library call = 42 launches (10 Adam steps); fp32 SIMT contraction 40x320x512 per "
                                 "stream, issue-bound, state L2-resident (DESIGN.md)",
    }
    tensor_peak, tensor_src = measured_tensor_peak()
    kern = {}
    for n, (c, tot, m) in summ.items():
        row = {"calls_per_step": c // reps, "mean_us": round(m, 2), "total_us_per_step": round(tot / reps, 1)}
        if n in tl.flops:
            row["achieved_tflops_fp32_equiv"] = round(tl.flops[n] / (tot * 1e-6) / 1e12, 2)
        else:
            row["achieved_gbs"] = round(alg_bytes.get(n, 0) / (m * 1e-6) / 1e9, 2)
        kern[n] = row
    top = max(summ.items(), key=lambda kv: kv[1][1])[0]
    if top in tl.flops:
        # tcgen05 3xTF32 GEMM: fp32-equivalent flops (2MNK per launch, summed over the launches of the step) against
        # the fp32-equivalent tensor peak = measured bf16 dense peak / 2 (tf32 rate) / 3 (three tf32 products per flop)
        ach = tl.flops[top] / (summ[top][1] * 1e-6) / 1e12
        roofline = {"bound": "tensor", "kernel": top + " (gemm_tf32x3_kernel)", "achieved": round(ach, 2),
                    "peak": round(tensor_peak, 1), "unit": "TFLOP/s", "frac": round(ach / tensor_peak, 4),
                    "traffic": GEMM_DRAM_TRAFFIC_BYTES, "peak_source": tensor_src,
                    "algorithmic_flops_per_step": tl.flops[top] // reps, "launches_per_step": summ[top][0] // reps,
                    "launch_us": round(summ[top][2], 2),
                    "note": "fp32-equivalent flops; every flop costs three tf32 tensor-core products (3xTF32 split)",
                    "kernels": kern}
    else:
        ach = alg_bytes.get(top, 0) / (summ[top][2] * 1e-6) / 1e9
        roofline = {"bound": "hbm", "kernel": top, "achieved": round(ach, 3), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 5), "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes.get(top, 0), "launch_us": round(summ[top][2], 2),
                    "note": notes.get(top, ""), "kernels": kern}
    # ---- the HBM-bound kernel of the path at the size where it is HBM-bound (cfg 4: K=1156, M=8, D=1024) ----------
    # (auxiliary measurements: a failure here must not cost the headline line)
    try:
        roofline["lvis_cache_step"] = lvis_cache_roofline(dev, peak, flush)
    except Exception as exc:      # noqa: BLE001
        roofline["lvis_cache_step"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    # ---- the tokenizer at the batch-64 shape of cfg 5 (the "FPS+kNN GB/s" part of BASELINE.json's metric) -------------
    try:
        roofline["tokenizer_cfg5"] = tokenizer_sweep(dev, peak, flush)
    except Exception as exc:      # noqa: BLE001
        roofline["tokenizer_cfg5"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    line = {"metric": METRIC, "value": round(value, 2), "unit": "clouds/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_resident / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams_per_gpu": S, "clouds_per_step": S * world,
                       "l2": "flushed before every timed step (256 MiB memset, then 256 MiB read so that no dirty lines remain)", "cuda_graph": not args.no_graph,
                       "parallelism": f"stream-per-GPU x{world} (no collective)"},
            "e2e": {"value": round(e2e, 2), "unit": "clouds/s", "h2d_bytes_per_step": S * N_POINTS * 12,
                    "d2h_bytes_per_step": S * N_CLASSES * 4, "ms_per_step": round(ms_e2e / K, 4)},
            "gpu_launches": launches * K * 2, "gpu_launches_per_step": launches,
            "roofline": roofline, "clocks": clocks.summary()}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, sec = cpu_reference_rate(5, 1, threads)
            line["cpu_baseline"] = {"value": round(rate, 4), "unit": "clouds/s", "cores": threads, "kind": "port",
                                    "sample": "5 sequential samples of one stream after 1 warm-up, same workload"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def lvis_cache_roofline(dev, peak, flush):
    """MODE-DOTA predict+fit at Objaverse-LVIS scale (K=1156, M=8, D=1024; 151 MB in+out per launch, > L2), beside a
    plain device copy of the same bytes under the same flush (the practical ceiling at this size)."""
    import uniadapter_b200 as ua
    from uniadapter_b200.streams import synthetic_text_features
    K, M, D = 1156, 8, 1024
    text = synthetic_text_features(K, D, seed=1).to(dev)
    model = ua.DOTA_mix(CFG, D, K, text.t().contiguous(), num_modes=M, device=dev)
    x = torch.nn.functional.normalize(torch.randn(1, D, device=dev), dim=-1)
    g = torch.softmax(100 * x @ text.t(), 1)
    src = torch.empty(2 * K * M * D, device=dev)
    dst = torch.empty_like(src)

    def median_us(fn, n=15):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        return sorted(ts)[len(ts) // 2]

    us = median_us(lambda: model.predict_then_fit(x, x, g))
    copy_us = median_us(lambda: dst.copy_(src))
    by = 16 * K * M * D
    return {"kernel": "ua_modedota_step_f32 (modedota_b1_kernel)", "workload": "K=1156 M=8 D=1024 predict+fit (cfg 4)",
            "bound": "hbm", "launch_us": round(us, 2), "algorithmic_bytes_per_launch": by,
            "achieved": round(by / us / 1e3, 1), "peak": peak, "unit": "GB/s", "frac": round(by / us / 1e3 / peak, 4),
            "traffic": LVIS_DRAM_TRAFFIC_BYTES,
            "plain_copy_same_bytes_us": round(copy_us, 2), "frac_of_plain_copy": round(copy_us / us, 4)}


def tokenizer_sweep(dev, peak, flush):
    """The tokenizer of BASELINE cfg 5 (Uni3D-L geometry: 64 clouds x 1024 xyz+rgb points, 512 groups x 64 neighbours):
    FPS + kNN grouping as achieved GB/s over SURVEY 8d's algorithmic bytes (cloud in, centres + groups out). FPS is a
    511-step dependent argmax chain per cloud (latency-bound by construction), so the fraction is reported, not a target."""
    import uniadapter_b200 as ua
    from uniadapter_b200.streams import unit_sphere_clouds
    B, N, G, k = 64, 1024, 512, 64
    g = torch.Generator().manual_seed(5)
    xyz = unit_sphere_clouds(B, N, g).to(dev)
    rgb = torch.rand(B, N, 3, generator=g).to(dev)
    _, centers = ua.fps_sample(xyz, G, None)

    def median_us(fn, n=9):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        return sorted(ts)[len(ts) // 2]

    fps_us = median_us(lambda: ua.fps_sample(xyz, G, None))
    knn_us = median_us(lambda: ua.knn_group(xyz, centers, k, rgb))
    by = B * (N * 6 * 4 + G * 12 + G * k * 6 * 4)
    tot = fps_us + knn_us
    return {"workload": "cfg 5 tokenizer: 64 clouds x 1024 xyz+rgb points, 512 groups x 64 neighbours",
            "kernels": "ua_fps_f32 (fps_reg_kernel) + ua_knn_group_f32 (knn_group_kernel)", "bound": "hbm",
            "fps_us": round(fps_us, 1), "knn_group_us": round(knn_us, 1), "algorithmic_bytes": by,
            "achieved": round(by / tot / 1e3, 1), "peak": peak, "unit": "GB/s", "frac": round(by / tot / 1e3 / peak, 4),
            "clouds_per_s": round(B / tot * 1e6),
            "note": "FPS is latency-bound (serial argmax chain, one CTA per cloud); kNN is issue-bound (selection)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--streams", type=int, default=15, help="corruption streams advanced in lock-step per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
