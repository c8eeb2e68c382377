"""Lock-step multi-stream engine: S independent corruption streams adapted side by side on one GPU.

The reference adapts one sample at a time, which leaves a B200 almost idle (the per-sample step is a chain of
latency-bound launches). Corruption streams are independent (fresh adapter per stream, Uni_Adapter.py:328-339), so
this engine advances S of them together: one tokenizer launch over S clouds, one encoder forward at batch S, one head
launch with S text matrices, one MODE-DOTA launch over the stacked state [S,K,M,D], one fusion launch. All buffers
are static and the no-grad part of the step is captured into a CUDA graph; inputs arrive from pinned host memory.

Per-stream semantics are exactly those of ``adapter.test_zeroshot_3d_core`` at batch 1 (the parity test runs both).
"""
from __future__ import annotations

import torch

from . import _lib
from .dota_mixture import init_state
from .fusion import fuse_logits
from .head import zero_shot_head
from .residual import ResidualLearner


class MultiStreamModeDota:
    """S MODE-DOTA adapters with stacked state: mu,var (S,K,M,D); pi,c (S,K,M); class_counts (S,K)."""

    def __init__(self, cfg, D, K, text, M, S, device):
        self.S, self.K, self.M, self.D = S, K, M, D
        self.epsilon = cfg.get('epsilon', 0.001)
        self.device = torch.device(device)
        text = text.to(self.device).float()
        if text.dim() == 2:
            text = text.unsqueeze(0).expand(S, -1, -1)
        states = [init_state(text[s].t().contiguous(), M, cfg.get('sigma', 1.0), self.device) for s in range(S)]
        self.mu = torch.stack([st[0] for st in states]).contiguous()
        self.var = torch.stack([st[1] for st in states]).contiguous()
        self.pi = torch.stack([st[2] for st in states]).contiguous()
        self.c = torch.stack([st[3] for st in states]).contiguous()
        self.class_counts = torch.stack([st[4] for st in states]).contiguous()
        self.t = 0

    def step(self, x_pred, x_fit, gamma_class, out_logits=None):
        """x_pred (S,Bp,D) | None, x_fit (S,B,D) | None, gamma_class (S,B,K) -> logits (S,Bp,K) | None."""
        S, K, M, D = self.S, self.K, self.M, self.D
        Bp = x_pred.shape[1] if x_pred is not None else 0
        B = x_fit.shape[1] if x_fit is not None else 0
        if Bp and out_logits is None:
            out_logits = torch.empty((S, Bp, K), dtype=torch.float32, device=self.device)
        rc = _lib.lib().ua_modedota_step_f32(
            _lib.ptr(x_pred), Bp, _lib.ptr(x_fit), _lib.ptr(gamma_class), B, K, 0, _lib.ptr(self.mu), _lib.ptr(self.var),
            _lib.ptr(self.pi), _lib.ptr(self.c), _lib.ptr(self.class_counts), S, K, M, D, float(self.epsilon),
            _lib.ptr(out_logits), K, 0, _lib.stream_ptr())
        _lib.check(rc, "ua_modedota_step_f32")
        self.t += B
        return out_logits

    def sample_step(self, x, x_aug, gamma_class, out_logits):
        """predict(x.half()) + fit(x) + fit(x_aug) of every stream in ONE pass over the stacked cache
        (csrc/modedota_sample.cu). x, x_aug (S,D); gamma_class (S,K); out_logits (S,K) written in place.
        Returns False when the shape is outside the kernel's register tiling (the caller then runs the two-pass form)."""
        S, K, M, D = self.S, self.K, self.M, self.D
        rc = _lib.lib().ua_modedota_sample_step_f32(
            _lib.ptr(x), _lib.ptr(x_aug), _lib.ptr(gamma_class), K, 0, _lib.ptr(self.mu), _lib.ptr(self.var),
            _lib.ptr(self.pi), _lib.ptr(self.c), _lib.ptr(self.class_counts), S, K, M, D, float(self.epsilon),
            _lib.ptr(out_logits), K, 0, _lib.stream_ptr())
        if rc == _lib.UA_ERR_UNSUPPORTED:
            return False
        _lib.check(rc, "ua_modedota_sample_step_f32")
        self.t += 2
        return True


class StreamEngine:
    """One adaptation step for S streams per call to :meth:`step`.

    encoder: a module from ``encoders.build_encoder`` (tokenizer inside); text: (K,D) unit rows shared by the
    streams at start. ``res_learning`` keeps one residual matrix and one Adam state per stream.
    """

    def __init__(self, encoder, vlm3d, text, num_streams, npoints, cfg, mode_M=8, res_learning=True, device='cuda',
                 use_graph=True, colored=False, seed=42, batch_views=True, stream_ids=None, external_rng=False):
        """``stream_ids``: global index of every local stream (default 0..S-1); stream s draws its jitter noise and FPS
        start indices from the counter-based generator keyed by ``seed + stream_ids[s]`` (csrc/rng.cu), so what a stream
        sees does not depend on which streams share its GPU or on the world size.
        ``external_rng``: the step reads start indices and noise from the static buffers ``start_buf`` (2,S) /
        ``noise_buf`` (S,N,3) instead, which the caller fills with :meth:`set_rng` before every step (parity harness:
        replays the draws of a CPU run of the reference, also under CUDA-graph replay)."""
        self.dev = torch.device(device)
        self.encoder, self.vlm3d = encoder, vlm3d
        self.S, self.N = num_streams, npoints
        self.cfg, self.M = cfg, mode_M
        self.text0 = text.to(self.dev).float().contiguous()
        self.K, self.D = self.text0.shape
        self.res_learning = res_learning
        self.use_graph = use_graph
        self.colored = colored
        self.batch_views = batch_views      # both views of a step through the encoder as one batch of 2S clouds
        self.fused_cache_pass = True        # predict + fit + fit in one cache pass (False: the two-launch sequence)
        S, N, K, D = self.S, self.N, self.K, self.D
        self.adapter = MultiStreamModeDota(cfg, D, K, self.text0, mode_M, S, self.dev)
        # static buffers
        self.pc = torch.zeros(S, N, 3, device=self.dev)
        self.rgb = torch.ones(S, N, 3, device=self.dev)
        if res_learning:   # residuals + Adam state per stream; learner.text = normalize(text0 + residual)
            self.learner = ResidualLearner(self.text0, S, mode_M, self.dev)
            self.text = self.learner.text
            self.residuals = self.learner.residual
        else:
            self.text = self.text0.unsqueeze(0).repeat(S, 1, 1).contiguous()    # (normalised) text per stream
        self.final = torch.zeros(S, K, device=self.dev)
        self.pred = torch.zeros(S, dtype=torch.int32, device=self.dev)
        self.dota_logits = torch.zeros(S, 1, K, device=self.dev)
        # random inputs of a step: static buffers, filled by csrc/rng.cu inside the step or by set_rng() before it
        self.external_rng = external_rng
        self.random_start = vlm3d in ('ulip', 'openshape')      # Uni3D samples from point 0 (pointnet2_ops)
        ids = list(range(S)) if stream_ids is None else [int(i) for i in stream_ids]
        if len(ids) != S:
            raise ValueError("stream_ids must name every local stream")
        self.stream_seeds = torch.tensor([seed + i for i in ids], dtype=torch.int64, device=self.dev)
        self.rng_step = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._rng_done = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.noise_buf = torch.zeros(S, N, 3, device=self.dev)
        self.start_buf = torch.zeros(2, S, dtype=torch.int64, device=self.dev)
        self.step_idx = 0
        self.graph = None
        self._host_out = torch.empty(S, K, dtype=torch.float32).pin_memory() if self.dev.type == 'cuda' else None

    # ---- pieces of the step ------------------------------------------------------------------------------------
    def _encode(self, pc, rgb=None):
        rgb = self.rgb if rgb is None else rgb
        if self.vlm3d == 'uni3d':
            return self.encoder.encode_pc(torch.cat((pc, rgb), dim=-1))
        if self.vlm3d == 'ulip':
            return self.encoder(pc)
        return self.encoder(pc, torch.cat((pc, rgb), dim=-1))

    def _set_start(self, start):
        if start is not None and self.random_start:
            for mod in self.encoder.modules():
                if hasattr(mod, 'next_start_idx'):
                    mod.next_start_idx = start

    def set_rng(self, start, start_aug, noise):
        """Parity harness (``external_rng=True``): the FPS start indices (S,) of the sample and of its jittered view and
        the N(0,1) noise (S,N,3) of the next step, copied into the static buffers the (captured) step reads."""
        if not self.external_rng:
            raise RuntimeError("set_rng needs StreamEngine(external_rng=True)")
        self.start_buf[0].copy_(start.reshape(-1), non_blocking=True)
        self.start_buf[1].copy_(start_aug.reshape(-1), non_blocking=True)
        self.noise_buf.copy_(noise, non_blocking=True)

    def _draw(self):
        """One launch: jitter noise and start indices of all S streams from their own counter-based generators."""
        rc = _lib.lib().ua_stream_rng_f32(_lib.ptr(self.stream_seeds), _lib.ptr(self.rng_step), self.S, self.N * 3,
                                          _lib.ptr(self.noise_buf), _lib.ptr(self.start_buf), self.N,
                                          _lib.ptr(self._rng_done), _lib.stream_ptr())
        _lib.check(rc, "ua_stream_rng_f32")

    @torch.no_grad()
    def _adapt(self):
        """Tokenizer + encoder, head, cache predict+fit, fit on the augmented view, fusion.

        The reference encodes the sample, adapts, then encodes the jittered copy (Uni_Adapter.py:401-431). The jittered
        cloud depends only on the input, so both views go through the tokenizer and the encoder as ONE batch of 2S
        clouds (per-cloud results are unchanged: every kernel of the pass is independent across clouds); the two cache
        steps keep the reference's order."""
        S, K = self.S, self.K
        if not self.external_rng:
            self._draw()
        pc2 = torch.cat((self.pc, self.pc + 0.05 * self.noise_buf), dim=0)          # Uni_Adapter.py:420-421
        if self.batch_views:
            self._set_start(self.start_buf.view(-1))
            emb = self._encode(pc2, torch.cat((self.rgb, self.rgb), dim=0))
            emb, emb_aug = emb[:S], emb[S:]
        else:
            self._set_start(self.start_buf[0])
            emb = self._encode(pc2[:S])
            self._set_start(self.start_buf[1])
            emb_aug = self._encode(pc2[S:])
        feats, clip_logits, _, prob, _ = zero_shot_head(emb, self.text)
        feats_aug, _, _, _, _ = zero_shot_head(emb_aug, self.text)
        # predict (fp16-rounded sample, Uni_Adapter.py:416) + fit + fit on the jittered view: one pass over the cache
        if not (self.fused_cache_pass and
                self.adapter.sample_step(feats, feats_aug, prob, self.dota_logits.view(S, K))):
            x_fit = feats.unsqueeze(1)                                     # (S,1,D): batch 1 per stream
            x_pred = x_fit.half().float()
            self.adapter.step(x_pred, x_fit, prob.unsqueeze(1), self.dota_logits)
            self.adapter.step(None, feats_aug.unsqueeze(1), prob.unsqueeze(1))
        self._clip_logits = clip_logits

    @torch.no_grad()
    def _fuse(self):
        final, arg, _ = fuse_logits(self._clip_logits, self.dota_logits.view(self.S, self.K), self.adapter.c,
                                    self.cfg['rho'], self.cfg['eta'], 1, 'mode_dota', per_row_c=True)
        self.final.copy_(final)
        self.pred.copy_(arg)

    def _learn_residuals(self):
        """10 alignment-loss backward / Adam rounds per stream (Uni_Adapter.py:443-476; the 11th loss evaluation of
        the reference has no effect on any state), all streams in one library call (csrc/residual.cu)."""
        a = self.adapter
        self.learner.learn(a.mu, a.var, a.pi, a.epsilon, iters=10)

    def _step_body(self, learn: bool):
        self._adapt()
        if learn:
            self._learn_residuals()
        self._fuse()

    # ---- public ------------------------------------------------------------------------------------------------
    def _run(self):
        learn = self.res_learning and self.step_idx > 0
        if self.use_graph and self.step_idx >= 2:
            if self.graph is None:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._step_body(learn)
            self.graph.replay()
        else:
            self._step_body(learn)
        self.step_idx += 1

    def step(self, pc_host: torch.Tensor, rgb_host: torch.Tensor | None = None, read_back: bool = True):
        """End-to-end step: pc_host (S,N,3) pinned host tensor -> (final_logits (S,K) pinned host, pred (S,) device).
        The host->device copy of the clouds and the device->host read of the fused logits are part of the step."""
        self.pc.copy_(pc_host, non_blocking=True)
        if rgb_host is not None:
            self.rgb.copy_(rgb_host, non_blocking=True)
        self._run()
        if read_back:
            self._host_out.copy_(self.final, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return self._host_out, self.pred
        return self.final, self.pred

    def step_device(self, pc_dev: torch.Tensor):
        """Same step with the clouds already resident in HBM; results stay on the device."""
        self.pc.copy_(pc_dev)
        self._run()
        return self.final, self.pred

    def launches_per_step(self) -> int:
        """Kernel launches of libua_b200.so in one (eager) step, counted by the library itself."""
        was_graph, self.use_graph = self.use_graph, False
        _lib.reset_launch_count()
        self._run()
        torch.cuda.synchronize()
        n = _lib.launch_count()
        self.use_graph = was_graph
        return n


class DotaEngine:
    """The DOTA branch (full covariance, ``--use-dota``, BASELINE cfg 1) of the per-sample loop as one CUDA-graph replay
    per sample: tokenizer + encoder, head, predict on the current Lambda, fit, update (the cooperative SPD inverse) and
    fusion, with static buffers. Same arithmetic as ``adapter.test_zeroshot_3d_core`` with ``use_dota``; the FPS start
    index comes from the device generator (graph-safe). One stream: the (K,D,D) covariance stack of DOTA is 42 MB per
    stream at cfg 1 and its kernels take one adapter per launch."""

    def __init__(self, encoder, vlm3d, text, npoints, cfg, device='cuda', use_graph=True, seed=42, stream_id=0,
                 external_rng=False, external_lambda=False):
        """``external_rng`` / ``external_lambda`` (parity harness): the step reads the FPS start index from
        ``start_buf`` (filled by :meth:`set_rng`) and, after its own ``update()``, overwrites Lambda with ``lambda_buf``
        (filled by :meth:`set_lambda` with the reference's Lambda of that step), so that the fp16 discriminant of the
        next sample is evaluated on the reference's precision matrix -- both as static buffers, so the
        captured graph is the path under test."""
        from .dota import DOTA
        self.dev = torch.device(device)
        self.encoder, self.vlm3d, self.cfg = encoder, vlm3d, cfg
        self.text = text.to(self.dev).float().contiguous()
        self.K, self.D = self.text.shape
        self.N = npoints
        self.adapter = DOTA(cfg, self.D, self.K, torch.full((self.D, self.K), 0.001), device=self.dev)   # Uni_Adapter.py:329-330
        self.pc = torch.zeros(1, npoints, 3, device=self.dev)
        self.rgb = torch.ones(1, npoints, 3, device=self.dev)
        self.final = torch.zeros(1, self.K, device=self.dev)
        self.pred = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._host_out = torch.empty(1, self.K, dtype=torch.float32).pin_memory()
        self.use_graph, self.graph, self.step_idx = use_graph, None, 0
        self.external_rng, self.external_lambda = external_rng, external_lambda
        self.random_start = vlm3d in ('ulip', 'openshape')
        self.stream_seeds = torch.tensor([seed + int(stream_id)], dtype=torch.int64, device=self.dev)
        self.rng_step = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._rng_done = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._rng_sink = torch.zeros(4, device=self.dev)
        self.start_buf = torch.zeros(2, 1, dtype=torch.int64, device=self.dev)
        self.lambda_buf = torch.zeros_like(self.adapter.Lambda) if external_lambda else None

    def set_rng(self, start):
        self.start_buf[0].copy_(start.reshape(-1), non_blocking=True)

    def set_lambda(self, lam):
        self.lambda_buf.copy_(lam, non_blocking=True)

    def _encode(self, pc):
        if self.random_start:
            if not self.external_rng:
                rc = _lib.lib().ua_stream_rng_f32(_lib.ptr(self.stream_seeds), _lib.ptr(self.rng_step), 1, 4,
                                                  _lib.ptr(self._rng_sink), _lib.ptr(self.start_buf), self.N,
                                                  _lib.ptr(self._rng_done), _lib.stream_ptr())
                _lib.check(rc, "ua_stream_rng_f32")
            for mod in self.encoder.modules():
                if hasattr(mod, 'next_start_idx'):
                    mod.next_start_idx = self.start_buf[0]
        if self.vlm3d == 'uni3d':
            return self.encoder.encode_pc(torch.cat((pc, self.rgb), dim=-1))
        if self.vlm3d == 'ulip':
            return self.encoder(pc)
        return self.encoder(pc, torch.cat((pc, self.rgb), dim=-1))

    @torch.no_grad()
    def _body(self):
        a, cfg = self.adapter, self.cfg
        feats, clip_logits, _, prob, _ = zero_shot_head(self._encode(self.pc), self.text)
        dl = a.predict(feats.mean(0).unsqueeze(0).half())
        a.fit(feats, prob)
        a.update()
        self.own_lambda = a.Lambda.clone() if self.external_lambda else None    # what update() produced (checked apart)
        if self.external_lambda:
            a.Lambda.copy_(self.lambda_buf)
        final, arg, _ = fuse_logits(clip_logits, dl, a.c, cfg['rho'], cfg['eta'], feats.shape[0], 'dota')
        self.final.copy_(final)
        self.pred.copy_(arg)

    def _run(self):
        if self.use_graph and self.step_idx >= 2:
            if self.graph is None:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._body()
            self.graph.replay()
        else:
            self._body()
        self.step_idx += 1

    def step(self, pc_host: torch.Tensor):
        """pc_host (1,N,3) pinned host tensor -> (final_logits (1,K) pinned host, pred (1,) device)."""
        self.pc.copy_(pc_host, non_blocking=True)
        self._run()
        self._host_out.copy_(self.final, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host_out, self.pred

    def step_device(self, pc_dev: torch.Tensor):
        """Same step with the cloud already resident in HBM; results stay on the device."""
        self.pc.copy_(pc_dev)
        self._run()
        return self.final, self.pred


class ShardedSampleEngine:
    """BASELINE cfg 4 as a product path: ONE stream whose MODE-DOTA cache is sharded by class over the ranks of the
    process group (Objaverse-LVIS: 1156 classes x 8 modes x 1024 dims = 76 MB of state, split into contiguous class
    ranges). Every rank runs the (replicated) tokenizer + encoder on the same sample and its jittered view, then the fused
    class-sharded step (``parallel.FusedShardedModeDota``: local zero-shot logits -> exchange over NVLink -> predict + two
    fits on the local classes -> exchange -> fusion, one kernel). The whole per-sample step is one CUDA-graph replay.
    With one rank (or no process group) the same code runs unsharded (P = 1).

    The jitter noise and the FPS start indices come from the stream's counter-based generator (``seed + stream_id``): every
    rank draws the same values without communicating."""

    def __init__(self, encoder, vlm3d, text, npoints, cfg, mode_M=8, device='cuda', use_graph=True, seed=42, stream_id=0,
                 emulate_world=None):
        from .parallel import FusedShardedModeDota
        import torch.distributed as dist
        self.dev = torch.device(device)
        self.encoder, self.vlm3d, self.cfg = encoder, vlm3d, cfg
        text = text.to(self.dev).float().contiguous()
        self.K, self.D = text.shape
        self.N = npoints
        if emulate_world is None and not (dist.is_available() and dist.is_initialized()):
            emulate_world = 1
        self.sharded = FusedShardedModeDota(cfg, text, mode_M, self.dev, emulate_world=emulate_world, use_graph=False)
        self.pc = torch.zeros(1, npoints, 3, device=self.dev)
        self.rgb = torch.ones(1, npoints, 3, device=self.dev)
        self.random_start = vlm3d in ('ulip', 'openshape')
        self.stream_seeds = torch.tensor([seed + int(stream_id)], dtype=torch.int64, device=self.dev)
        self.rng_step = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._rng_done = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.noise_buf = torch.zeros(1, npoints, 3, device=self.dev)
        self.start_buf = torch.zeros(2, 1, dtype=torch.int64, device=self.dev)
        self.use_graph, self.graph, self.step_idx = use_graph, None, 0
        self._host_out = torch.empty(1, self.K, dtype=torch.float32).pin_memory()

    @property
    def final(self):
        return self.sharded.mine.out_final

    @property
    def pred(self):
        return self.sharded.mine.out_argmax

    @torch.no_grad()
    def _body(self):
        rc = _lib.lib().ua_stream_rng_f32(_lib.ptr(self.stream_seeds), _lib.ptr(self.rng_step), 1, self.N * 3,
                                          _lib.ptr(self.noise_buf), _lib.ptr(self.start_buf), self.N,
                                          _lib.ptr(self._rng_done), _lib.stream_ptr())
        _lib.check(rc, "ua_stream_rng_f32")
        pc2 = torch.cat((self.pc, self.pc + 0.05 * self.noise_buf), dim=0)          # Uni_Adapter.py:420-421
        rgb2 = torch.cat((self.rgb, self.rgb), dim=0)
        if self.random_start:
            for mod in self.encoder.modules():
                if hasattr(mod, 'next_start_idx'):
                    mod.next_start_idx = self.start_buf.view(-1)
        if self.vlm3d == 'uni3d':
            emb = self.encoder.encode_pc(torch.cat((pc2, rgb2), dim=-1))
        elif self.vlm3d == 'ulip':
            emb = self.encoder(pc2)
        else:
            emb = self.encoder(pc2, torch.cat((pc2, rgb2), dim=-1))
        self.sharded.enqueue(emb)

    def _run(self):
        if self.use_graph and self.step_idx >= 2:
            if self.graph is None:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._body()
            self.graph.replay()
        else:
            self._body()
        self.step_idx += 1

    def step(self, pc_host: torch.Tensor, rgb_host: torch.Tensor | None = None):
        """pc_host (1,N,3) pinned host tensor -> (final_logits (1,K) pinned host, pred (1,) device); copies included."""
        self.pc.copy_(pc_host, non_blocking=True)
        if rgb_host is not None:
            self.rgb.copy_(rgb_host, non_blocking=True)
        self._run()
        self._host_out.copy_(self.final, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host_out, self.pred

    def step_device(self, pc_dev: torch.Tensor):
        self.pc.copy_(pc_dev)
        self._run()
        return self.final, self.pred

    def check(self):
        self.sharded.check()
