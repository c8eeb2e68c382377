"""Multi-GPU partitioning — only where the path shards. One process per GPU, torch.distributed (NCCL).

1. **Stream per GPU** (BASELINE cfg 2): corruption streams are independent (a fresh adapter per stream,
   Uni_Adapter.py:328-339), so stream ``s`` runs on rank ``s mod P`` with no data-path communication; accuracies are
   gathered once at the end. Per-stream RNG seeds (``seed + stream``) make a stream's result independent of P.
2. **Class-sharded MODE-DOTA cache** (cfg 4, Objaverse-LVIS, K=1156): ``mu,var,pi,c,class_counts`` and the text rows are
   split by contiguous class ranges. Every rank sees the same sample (replicated encoder), computes its slice of the
   zero-shot and cache logits, ONE all-gather of ``2*K_pad`` floats per rank assembles both rows, then softmax / fusion
   are replicated and ``fit`` is purely local. ``c.mean()`` comes from the closed form ``K + fits*B`` (every fit adds
   exactly ``sum_b sum_k gamma_class = B``), so no second collective is needed.

The collective and the host logic are backend-agnostic (``gloo`` on CPU in the tests); the per-shard compute goes through
``CudaShardOps`` (libua_b200.so) in the product.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


# ----------------------------------------------------------------------------------------------------------
# partitioning arithmetic
# ----------------------------------------------------------------------------------------------------------
def assign_streams(num_streams: int, world: int, rank: int) -> list[int]:
    """Streams of this rank: s -> rank s mod P (15 streams on 8 GPUs = two waves, 8 + 7)."""
    return [s for s in range(num_streams) if s % world == rank]


def class_partition(K: int, world: int) -> list[tuple[int, int]]:
    """Contiguous class ranges; the first K mod P ranks own one class more (1156 / 8 -> 4 x 145 + 4 x 144)."""
    base, extra = divmod(K, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def padded_shard(K: int, world: int) -> int:
    return -(-K // world)


def closed_form_count_sum(K: int, fits: int, batch: int) -> float:
    """sum(c) of a MODE-DOTA cache after ``fits`` fit calls of batch ``batch`` (initial counts sum to K)."""
    return float(K + fits * batch)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ----------------------------------------------------------------------------------------------------------
# 1. stream per GPU
# ----------------------------------------------------------------------------------------------------------
def gather_stream_results(local: dict[int, dict], num_streams: int) -> dict[int, dict] | None:
    """local: {stream index: {'acc1','acc3','acc5'}} of this rank -> all streams on rank 0 (None elsewhere).
    One small collective per run (3 floats per stream), never per step."""
    rank, world = world_info()
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.is_initialized() and dist.get_backend() == 'nccl' \
        else torch.device('cpu')
    table = torch.full((num_streams, 3), float('nan'), dtype=torch.float64, device=dev)
    for s, r in local.items():
        table[s] = torch.tensor([r['acc1'], r['acc3'], r['acc5']], dtype=torch.float64)
    if world > 1:
        parts = [torch.empty_like(table) for _ in range(world)]
        dist.all_gather(parts, table)
        stacked = torch.stack(parts)                       # (P, streams, 3): exactly one rank holds a non-NaN row
        table = torch.nan_to_num(stacked, nan=0.0).sum(0)
        owned = (~torch.isnan(stacked[..., 0])).sum(0)
        if not bool((owned == 1).all()):
            raise RuntimeError("stream partition is not a partition: some stream has no / several owners")
    if rank != 0:
        return None
    return {s: dict(acc1=float(table[s, 0]), acc3=float(table[s, 1]), acc5=float(table[s, 2])) for s in range(num_streams)}


# ----------------------------------------------------------------------------------------------------------
# 2. class-sharded cache
# ----------------------------------------------------------------------------------------------------------
class CudaShardOps:
    """Per-shard compute on the GPU (libua_b200.so). State: MODE-DOTA over this rank's K_p classes."""

    def __init__(self, cfg, D, text_shard, M, device):
        from .engine import MultiStreamModeDota
        self.dev = torch.device(device)
        self.text = text_shard.to(self.dev).float().contiguous()        # (K_p, D)
        self.Kp = self.text.shape[0]
        self.cache = MultiStreamModeDota(cfg, D, self.Kp, self.text, M, 1, self.dev)

    def head_local(self, feats_raw, out_row):
        """xnorm (B,D) of the raw features; out_row[0:K_p] <- 100 * xnorm[0] @ text_shard^T (row of the gather buffer)."""
        from . import _lib
        feats_raw = feats_raw.float().contiguous()
        B, D = feats_raw.shape
        xn = torch.empty_like(feats_raw)
        lg = out_row if B == 1 else torch.empty((B, self.Kp), dtype=torch.float32, device=self.dev)
        rc = _lib.lib().ua_head_f32(_lib.ptr(feats_raw), B, D, _lib.ptr(self.text), 1, self.Kp, 100.0, _lib.ptr(xn),
                                    _lib.ptr(lg), None, None, None, _lib.stream_ptr())
        _lib.check(rc, "ua_head_f32")
        return xn

    def predict_local(self, x_pred, out_row):
        """out_row[0:K_p] <- cache predict(x_pred) on the current state, written straight into the gather buffer."""
        from . import _lib
        c = self.cache
        rc = _lib.lib().ua_modedota_step_f32(
            _lib.ptr(x_pred.contiguous()), 1, None, None, 0, 0, 0, _lib.ptr(c.mu), _lib.ptr(c.var), _lib.ptr(c.pi),
            _lib.ptr(c.c), _lib.ptr(c.class_counts), 1, self.Kp, c.M, c.D, float(c.epsilon), _lib.ptr(out_row),
            out_row.shape[0], 0, _lib.stream_ptr())
        _lib.check(rc, "ua_modedota_step_f32")

    def fit(self, x, prob_full, k_lo):
        """fit on this shard with columns [k_lo, k_lo+K_p) of the replicated prob_map (B, K)."""
        from . import _lib
        c = self.cache
        B = x.shape[0]
        rc = _lib.lib().ua_modedota_step_f32(
            None, 0, _lib.ptr(x.contiguous()), _lib.ptr(prob_full.contiguous()), B, prob_full.shape[1], k_lo,
            _lib.ptr(c.mu), _lib.ptr(c.var), _lib.ptr(c.pi), _lib.ptr(c.c), _lib.ptr(c.class_counts), 1, self.Kp, c.M,
            c.D, float(c.epsilon), None, 0, 0, _lib.stream_ptr())
        _lib.check(rc, "ua_modedota_step_f32")

    def fuse(self, clip, dota, c_sum, c_count, rho, eta, batch):
        """c_sum: the closed-form sum of the soft counts, a float or a 1-element device tensor (CUDA-graph safe)."""
        from .fusion import fuse_logits
        if torch.is_tensor(c_sum):
            final, arg, _ = fuse_logits(clip, dota, c_sum.view(1), rho, eta, batch, 'mode_dota', c_count=c_count)
        else:
            final, arg, _ = fuse_logits(clip, dota, None, rho, eta, batch, 'mode_dota', c_sum=c_sum, c_count=c_count)
        return final, arg

    def softmax(self, logits):
        return torch.softmax(logits, dim=1)

    def empty(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.dev)


@dataclass
class ShardedStepOutput:
    final_logits: torch.Tensor     # (1, K) replicated
    pred: int
    clip_logits: torch.Tensor
    dota_logits: torch.Tensor


class ShardedModeDota:
    """Class-sharded MODE-DOTA step for batch-1 samples (cfg 4). ``ops_factory(text_shard)`` builds the per-shard
    compute object (``CudaShardOps`` in the product; the CPU tests inject an oracle-backed one)."""

    def __init__(self, cfg, text, M, ops_factory, group=None, rank=None, world=None, gather_fn=None):
        self.cfg, self.M = cfg, M
        self.K, self.D = text.shape
        self.group = group
        self.rank, self.world = world_info()
        if rank is not None:            # explicit placement: single-process emulation of P ranks (tests, 1-GPU boxes)
            self.rank, self.world = rank, world
        self.gather_fn = gather_fn
        self.ranges = class_partition(self.K, self.world)
        self.k_lo, self.k_hi = self.ranges[self.rank]
        self.K_pad = padded_shard(self.K, self.world)
        self.ops = ops_factory(text[self.k_lo:self.k_hi])
        self.fits = 0
        self.send = self.ops.empty(2, self.K_pad)
        self.scratch_row = self.ops.empty(self.K_pad)
        self.recv = self.ops.empty(self.world, 2, self.K_pad)
        self._graph = None

    def _all_gather(self):
        if self.gather_fn is not None:
            self.gather_fn(self)
            return
        if self.world == 1:
            self.recv[0].copy_(self.send)
            return
        if dist.get_backend(self.group) == 'nccl':
            dist.all_gather_into_tensor(self.recv.view(-1), self.send.view(-1), group=self.group)
        else:
            parts = list(self.recv.unbind(0))
            dist.all_gather(parts, self.send, group=self.group)

    def assemble(self, recv):
        """(P, 2, K_pad) gathered buffer -> clip (1,K), dota (1,K) in class order (drops the padding)."""
        clip = torch.cat([recv[r, 0, :hi - lo] for r, (lo, hi) in enumerate(self.ranges)]).unsqueeze(0)
        dota = torch.cat([recv[r, 1, :hi - lo] for r, (lo, hi) in enumerate(self.ranges)]).unsqueeze(0)
        return clip, dota

    def step(self, feats_raw: torch.Tensor, feats_aug_raw: torch.Tensor | None) -> ShardedStepOutput:
        """feats_raw / feats_aug_raw: (1, D) raw encoder outputs of the sample and of its augmented view (replicated)."""
        self.local_logits(feats_raw)
        self._all_gather()                                              # the one exchange of the step
        return self.finish(feats_aug_raw)

    def local_logits(self, feats_raw):
        """Phase 1 (before the exchange): local zero-shot and cache logits into the send buffer."""
        ops = self.ops
        self._x = ops.head_local(feats_raw, self.send[0])               # xnorm; local zero-shot logits -> send[0]
        x_pred = self._x.mean(0, keepdim=True).half().float()           # Uni_Adapter.py:416
        ops.predict_local(x_pred, self.send[1])                         # local cache logits -> send[1]

    def finish(self, feats_aug_raw, device_counts=None):
        """Phase 2 (after the exchange): replicated softmax / fusion, local fits. ``device_counts``: 1-element device
        tensor holding sum(c) (advanced on the device, for CUDA-graph replay); the prediction then stays a tensor."""
        cfg, ops, x = self.cfg, self.ops, self._x
        clip, dota = self.assemble(self.recv)
        prob = ops.softmax(clip)
        ops.fit(x, prob, self.k_lo)
        self.fits += 1
        if feats_aug_raw is not None:
            x_aug = ops.head_local(feats_aug_raw, self.scratch_row)     # only its xnorm is used
            ops.fit(x_aug, prob, self.k_lo)                             # augmented view, original prob_map (:430)
            self.fits += 1
        if device_counts is not None:
            device_counts.add_(float(x.shape[0] * (2 if feats_aug_raw is not None else 1)))
            final, arg = ops.fuse(clip, dota, device_counts, self.K * self.M, cfg['rho'], cfg['eta'], x.shape[0])
            return ShardedStepOutput(final, arg, clip, dota)
        c_sum = closed_form_count_sum(self.K, self.fits, x.shape[0])
        final, arg = ops.fuse(clip, dota, c_sum, self.K * self.M, cfg['rho'], cfg['eta'], x.shape[0])
        return ShardedStepOutput(final, int(arg[0]), clip, dota)

    def step_graphed(self, feats_raw: torch.Tensor, feats_aug_raw: torch.Tensor) -> ShardedStepOutput:
        """The same step as two CUDA-graph replays around the one exchange: graph A = local head + cache logits into the
        send buffer, then the all-gather (eager: NCCL stays outside the graphs), graph B = replicated softmax / fusion and
        the local fits. No host work between the ~15 launches of a step, which is what bounds the eager step (0.3 ms for
        a 33 us cache pass). Call after at least one eager ``step`` (communicator warm-up). ``pred`` of the result is a
        1-element device tensor; inputs are copied into static buffers, outputs are static buffers overwritten by the
        next replay. The product path is :class:`FusedShardedModeDota` (one kernel per step, exchange inside); this class
        keeps the collective-library form for comparison and for the CPU (gloo) tests of the host logic."""
        if self._graph is None:
            self._g_in = feats_raw.clone().contiguous()
            self._g_aug = feats_aug_raw.clone().contiguous()
            # sum(c) so far in closed form (H7): K initial counts + one per fitted row
            self._g_counts = torch.full((1,), closed_form_count_sum(self.K, self.fits, feats_raw.shape[0]),
                                        dtype=torch.float32, device=feats_raw.device)
            fits_before = self.fits
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                self.local_logits(self._g_in)
            with torch.cuda.graph(gb, pool=ga.pool()):   # graph B reads graph A's xnorm: one memory pool
                self._g_out = self.finish(self._g_aug, device_counts=self._g_counts)
            self._graph = (ga, gb)
            self.fits = fits_before       # capture runs no kernel: the counters advance on replay
        self._g_in.copy_(feats_raw)
        self._g_aug.copy_(feats_aug_raw)
        self._graph[0].replay()
        self._all_gather()                                              # the one exchange of the step (NCCL, eager)
        self._graph[1].replay()
        self.fits += 2
        return self._g_out


# ----------------------------------------------------------------------------------------------------------
# 2b. class-sharded cache, product path: one fused kernel per rank and step (csrc/modedota_sample.cu)
# ----------------------------------------------------------------------------------------------------------
class _FusedRank:
    """Device state of one rank: text / cache shard, static inputs, symmetric receive + flag buffers, counters, outputs."""

    def __init__(self, cfg, text, M, P, rank, dev, recv=None, flag=None):
        from .engine import MultiStreamModeDota
        K, D = text.shape
        self.rank = rank
        self.k_lo, self.k_hi = class_partition(K, P)[rank]
        self.Kp, self.K_pad = self.k_hi - self.k_lo, padded_shard(K, P)
        self.text = text[self.k_lo:self.k_hi].to(dev).float().contiguous()
        self.cache = MultiStreamModeDota(cfg, D, self.Kp, self.text, M, 1, dev)
        self.x2 = torch.zeros(2, D, device=dev)                      # raw features: sample, jittered view
        self.xn = torch.zeros(2, D, device=dev)                      # normalised
        self.clip2 = torch.zeros(2, self.Kp, device=dev)             # local zero-shot logits of both rows (row 0 is used)
        self.recv = recv if recv is not None else torch.zeros(2 * P * 2 * self.K_pad, device=dev)
        self.flag = flag if flag is not None else torch.zeros(2 * P, dtype=torch.int32, device=dev)
        self.seq = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.done = torch.zeros(2, dtype=torch.int32, device=dev)
        self.c_sum = torch.full((1,), float(K), dtype=torch.float32, device=dev)     # initial soft counts sum to K
        self.out_final = torch.zeros(1, K, device=dev)
        self.out_argmax = torch.zeros(1, dtype=torch.int32, device=dev)
        self.out_clip = torch.zeros(1, K, device=dev)
        self.out_dota = torch.zeros(1, K, device=dev)

    def struct(self, recv_ptrs, flag_ptrs, in_kernel_head=True):
        from ._lib import ShardRank
        c = self.cache
        if in_kernel_head:      # raw rows in, text rows in: the kernel normalises and forms the zero-shot logits itself
            return ShardRank(self.x2[0].data_ptr(), self.x2[1].data_ptr(), self.text.data_ptr(), None, c.mu.data_ptr(),
                             c.var.data_ptr(), c.pi.data_ptr(), c.c.data_ptr(), c.class_counts.data_ptr(), recv_ptrs.data_ptr(),
                             flag_ptrs.data_ptr(), self.seq.data_ptr(), self.err.data_ptr(), self.done.data_ptr(),
                             self.c_sum.data_ptr(), self.out_final.data_ptr(), self.out_argmax.data_ptr(),
                             self.out_clip.data_ptr(), self.out_dota.data_ptr(), self.rank, 0)
        return ShardRank(self.xn[0].data_ptr(), self.xn[1].data_ptr(), None, self.clip2.data_ptr(), c.mu.data_ptr(), c.var.data_ptr(),
                         c.pi.data_ptr(), c.c.data_ptr(), c.class_counts.data_ptr(), recv_ptrs.data_ptr(), flag_ptrs.data_ptr(),
                         self.seq.data_ptr(), self.err.data_ptr(), self.done.data_ptr(), self.c_sum.data_ptr(),
                         self.out_final.data_ptr(), self.out_argmax.data_ptr(), self.out_clip.data_ptr(),
                         self.out_dota.data_ptr(), self.rank, 0)


class FusedShardedModeDota:
    """Class-sharded MODE-DOTA sample step (BASELINE cfg 4) as ONE persistent kernel per rank and step:
    ``ua_head_f32`` (normalise the sample and its jittered view, local zero-shot logits) followed by
    ``ua_modedota_sharded_step_f32`` -- zero-shot logits pushed to every peer over NVLink, gathered prob_map, predict +
    fit + fit over the local classes with every cache logit stored straight into the peers, second flag exchange,
    fusion. No collective call, no host work between launches; ``step`` replays one CUDA graph per sample.

    Real ranks: one process per GPU (``torch.distributed`` initialised, NCCL); the receive and flag buffers come from
    ``torch.distributed._symmetric_memory`` so that every rank's buffers are mapped into every process.
    ``emulate_world=P``: all P ranks on ONE device, stepped by one cooperative launch (single-GPU tests);
    ``emulate_only=r`` then launches rank r alone (its peers never arrive: the time-out path)."""

    def __init__(self, cfg, text, M, device, group=None, emulate_world: int | None = None, use_graph: bool = True,
                 emulate_only: int | None = None, in_kernel_head: bool = True):
        from . import _lib
        import ctypes as C
        self.cfg, self.M = cfg, M
        self.K, self.D = text.shape
        self.dev = torch.device(device)
        self.use_graph, self._graph, self.steps = use_graph, None, 0
        self.emulated = emulate_world is not None
        if self.emulated:
            self.P = int(emulate_world)
            self.ranks = [_FusedRank(cfg, text, M, self.P, r, self.dev) for r in range(self.P)]
            recvs, flags = [r.recv for r in self.ranks], [r.flag for r in self.ranks]
            self._recv_ptrs = torch.tensor([t.data_ptr() for t in recvs], dtype=torch.int64, device=self.dev)
            self._flag_ptrs = torch.tensor([t.data_ptr() for t in flags], dtype=torch.int64, device=self.dev)
            if emulate_only is not None:      # tests: launch ONE of the emulated ranks, its peers never arrive
                self.ranks = [self.ranks[emulate_only]]
        else:
            import torch.distributed._symmetric_memory as symm_mem
            self.rank, self.P = world_info()
            group = group if group is not None else dist.group.WORLD
            K_pad = padded_shard(self.K, self.P)
            recv = symm_mem.empty(2 * self.P * 2 * K_pad, dtype=torch.float32, device=self.dev)
            flag = symm_mem.empty(2 * self.P, dtype=torch.int32, device=self.dev)
            recv.zero_()
            flag.zero_()
            h_recv, h_flag = symm_mem.rendezvous(recv, group), symm_mem.rendezvous(flag, group)
            self._handles = (h_recv, h_flag)
            self._recv_ptrs = torch.tensor(list(h_recv.buffer_ptrs), dtype=torch.int64, device=self.dev)
            self._flag_ptrs = torch.tensor(list(h_flag.buffer_ptrs), dtype=torch.int64, device=self.dev)
            self.ranks = [_FusedRank(cfg, text, M, self.P, self.rank, self.dev, recv, flag)]
            torch.cuda.synchronize()
            dist.barrier(group=group)      # every rank has zeroed its flags before anybody signals
        self.K_pad = self.ranks[0].K_pad
        # True: ONE launch per step (normalisation and zero-shot logits inside the kernel); False: ua_head_f32 first
        self.in_kernel_head = in_kernel_head
        self._structs = (_lib.ShardRank * len(self.ranks))(*[r.struct(self._recv_ptrs, self._flag_ptrs, in_kernel_head)
                                                              for r in self.ranks])
        self._structs_ptr = C.cast(self._structs, C.c_void_p)

    @property
    def mine(self) -> _FusedRank:
        return self.ranks[0]

    def _launch(self):
        from . import _lib
        lib = _lib.lib()
        for r in ([] if self.in_kernel_head else self.ranks):
            rc = lib.ua_head_f32(_lib.ptr(r.x2), 2, self.D, _lib.ptr(r.text), 1, r.Kp, 100.0, _lib.ptr(r.xn), _lib.ptr(r.clip2),
                                 None, None, None, _lib.stream_ptr())
            _lib.check(rc, "ua_head_f32")
        rc = lib.ua_modedota_sharded_step_f32(self._structs_ptr, len(self.ranks), self.P, self.K, self.K_pad, self.M, self.D,
                                              float(self.cfg.get('epsilon', 0.001)), float(self.cfg['rho']),
                                              float(self.cfg['eta']), _lib.stream_ptr())
        _lib.check(rc, "ua_modedota_sharded_step_f32")

    @torch.no_grad()
    def enqueue(self, feats2: torch.Tensor) -> ShardedStepOutput:
        """Enqueue one step on the current stream without any graph handling of its own (for callers that capture a
        larger graph around it, ``engine.ShardedSampleEngine``): feats2 (2,D) = raw features of the sample and of its view."""
        for r in self.ranks:
            r.x2.copy_(feats2.reshape(2, -1), non_blocking=True)
        self._launch()
        self.steps += 1
        m = self.mine
        return ShardedStepOutput(m.out_final, m.out_argmax, m.out_clip, m.out_dota)

    @torch.no_grad()
    def step(self, feats_raw: torch.Tensor, feats_aug_raw: torch.Tensor) -> ShardedStepOutput:
        """feats_raw / feats_aug_raw (1,D): raw encoder outputs of the sample and of its jittered view (the same on every
        rank). Returns the replicated result of this rank (static buffers, overwritten by the next step)."""
        for r in self.ranks:
            r.x2[0].copy_(feats_raw.reshape(-1), non_blocking=True)
            r.x2[1].copy_(feats_aug_raw.reshape(-1), non_blocking=True)
        if self.use_graph and self.steps >= 1:
            if self._graph is None:
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._launch()
            self._graph.replay()
        else:
            self._launch()
        self.steps += 1
        m = self.mine
        return ShardedStepOutput(m.out_final, m.out_argmax, m.out_clip, m.out_dota)

    def check(self):
        """Raise if any in-kernel wait for a peer ran out (host synchronisation: call every N steps and at stream end).
        After an error the outputs of that step are NaN / -1 and the cache shard of the waiting rank is untouched."""
        errs = [int(r.err.item()) for r in self.ranks]
        if any(errs):
            what = {1: "a peer's zero-shot logits did not arrive", 2: "a peer's cache logits did not arrive (or the peer aborted)"}
            raise RuntimeError("class-sharded step: " + "; ".join(f"rank {r.rank}: {what.get(e, e)}"
                                                                   for r, e in zip(self.ranks, errs) if e))
