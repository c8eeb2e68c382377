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

    def enable_p2p(self):
        """Replace the NCCL all-gather of the step by the peer-memory exchange kernel (csrc/p2p.cu): symmetric receive
        and flag buffers from torch.distributed._symmetric_memory (mapped into every rank's address space), one CTA per
        rank pushes its logits into every peer's buffer over NVLink and waits for the peers' flags. With it the whole
        step (local logits -> exchange -> replicated softmax / fusion -> local fits) is ONE CUDA graph."""
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        dev = self.send.device
        n, P = self.send.numel(), self.world
        group = self.group if self.group is not None else dist.group.WORLD
        self._sym_recv = symm_mem.empty(2 * P * n, dtype=torch.float32, device=dev)
        self._sym_flag = symm_mem.empty(P, dtype=torch.int32, device=dev)
        self._sym_recv.zero_()
        self._sym_flag.zero_()
        h_recv = symm_mem.rendezvous(self._sym_recv, group)
        h_flag = symm_mem.rendezvous(self._sym_flag, group)
        self._p2p_recv_ptrs = torch.tensor(list(h_recv.buffer_ptrs), dtype=torch.int64, device=dev)
        self._p2p_flag_ptrs = torch.tensor(list(h_flag.buffer_ptrs), dtype=torch.int64, device=dev)
        self._p2p_seq = torch.zeros(1, dtype=torch.int32, device=dev)
        self._p2p_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._p2p_handles = (h_recv, h_flag)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)          # every rank has zeroed its flags before anybody signals
        self._p2p = True

    def _all_gather(self):
        if self.gather_fn is not None:
            self.gather_fn(self)
            return
        if getattr(self, '_p2p', False):
            from . import _lib
            rc = _lib.lib().ua_p2p_allgather_f32(_lib.ptr(self.send), self.send.numel(), _lib.ptr(self._p2p_recv_ptrs),
                                                 _lib.ptr(self._p2p_flag_ptrs), self.rank, self.world,
                                                 _lib.ptr(self._p2p_seq), _lib.ptr(self.recv), _lib.ptr(self._p2p_err),
                                                 _lib.stream_ptr())
            _lib.check(rc, "ua_p2p_allgather_f32")
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
        next replay. (Capturing the NCCL all-gather INSIDE one graph did not complete in the one 2-GPU attempt of round 1;
        the planned replacement for the eager collective is a peer-memory exchange kernel.)"""
        if self._graph is None:
            self._g_in = feats_raw.clone().contiguous()
            self._g_aug = feats_aug_raw.clone().contiguous()
            # sum(c) so far in closed form (H7): K initial counts + one per fitted row
            self._g_counts = torch.full((1,), closed_form_count_sum(self.K, self.fits, feats_raw.shape[0]),
                                        dtype=torch.float32, device=feats_raw.device)
            fits_before = self.fits
            ga = torch.cuda.CUDAGraph()
            if getattr(self, '_p2p', False):                 # peer-memory exchange: the whole step is one graph
                with torch.cuda.graph(ga):
                    self.local_logits(self._g_in)
                    self._all_gather()
                    self._g_out = self.finish(self._g_aug, device_counts=self._g_counts)
                self._graph = (ga, None)
            else:
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga):
                    self.local_logits(self._g_in)
                with torch.cuda.graph(gb, pool=ga.pool()):   # graph B reads graph A's xnorm: one memory pool
                    self._g_out = self.finish(self._g_aug, device_counts=self._g_counts)
                self._graph = (ga, gb)
            self.fits = fits_before       # capture runs no kernel: the counters advance on replay
        self._g_in.copy_(feats_raw)
        self._g_aug.copy_(feats_aug_raw)
        self._graph[0].replay()
        if self._graph[1] is not None:
            self._all_gather()                                          # the one exchange of the step (NCCL, eager)
            self._graph[1].replay()
        self.fits += 2
        return self._g_out
