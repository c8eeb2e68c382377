"""TEST INFRASTRUCTURE — the reference's per-sample test-time loop driven through the reference's OWN modules.

``test_zeroshot_3d_core`` (Uni_Adapter.py:272-595) hard-codes CUDA events, ``.cuda()`` and a plotly dump, so it cannot be
called on the CPU; ``ReferenceStream.step`` restates only that scaffolding (Uni_Adapter.py:382-521, batch 1). Every
numerical call is the reference's: its encoder modules (models/ulip/pointbert/point_encoder.py PointTransformer,
models/openshape/ppta.py Projected / PointPatchTransformer), ``get_logits_wrapper`` and ``softmax_entropy``
(Uni_Adapter.py:21-26,53-75), ``DOTA`` (dota.py:19-87), ``DOTA_mix`` (dota_mixture.py:7-274),
``compute_text_alignment_loss`` (Uni_Adapter.py:191-270) and torch's Adam (Uni_Adapter.py:346-352,455-476).

Used by oracle/make_golden.py (golden vectors) and by ``bench.py --impl reference`` / ``cpu_baseline`` (the reference's
CPU implementation timed on the GPU box's host cores, ``kind: "reference"``). The reference sources are imported from
/root/reference when present (build container) or from the git-ignored copy oracle/_ref/ made by oracle/fetch_ref.py
(the GPU box). Never on the product path.
"""
from __future__ import annotations

import contextlib
import types

import torch
import torch.nn.functional as F

from . import reference_loader as R


@contextlib.contextmanager
def cpu_only():
    """The reference picks its device with ``torch.cuda.is_available()`` (dota.py:24, dota_mixture.py:31): on the GPU box the
    CPU arm must hide the GPU from it, or the adapter state lands on cuda:0 while the features are on the host."""
    orig = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        yield
    finally:
        torch.cuda.is_available = orig


def ulip_reference_model(depth: int, seed: int):
    """ULIP-2 point branch exactly as the reference composes it: PointTransformer trunk (768-d) @ pc_projection
    (models/ulip/ulip_model.py:13-18), random init under ``torch.manual_seed(seed)``. Returns (callable, trunk, proj)."""
    _, _, penc = R.ulip_pointbert()
    margs = types.SimpleNamespace(pc_feat_dim=768, pc_depth=depth, drop_path_rate=0.0, num_head=6, group_size=32,
                                  num_group=512, encoder_dim=256)
    torch.manual_seed(seed)
    trunk = penc.PointTransformer(margs)                       # models/ulip/pointbert/point_encoder.py:103
    proj = torch.empty(768, 512)
    torch.nn.init.normal_(proj, std=768 ** -0.5)
    trunk.eval()
    return (lambda xyz: trunk(xyz) @ proj), trunk, proj


def openshape_reference_model(depth: int, patches: int, seed: int):
    """OpenShape PointBERT scaling 4 (models/openshape/ppta.py:181-186) at ``depth`` transformer layers."""
    _, ppta = R.openshape_ppta()
    torch.manual_seed(seed)
    return ppta.Projected('global', ppta.PointPatchTransformer('global', None, 512, depth, 8, 512 * 3, 256, patches, 0.2,
                                                               64, 6), torch.nn.Linear(512, 1280)).eval()


class ReferenceStream:
    """One corruption stream adapted by the reference's code on the CPU, one sample per ``step``."""

    def __init__(self, model, vlm3d, text, cfg, feat_dim, mode_M=8, res_learning=False):
        self.ua = R.uni_adapter()
        self.model, self.cfg, self.text = model, cfg, text
        self.args = types.SimpleNamespace(vlm3d=vlm3d)
        K = text.shape[0]
        self.use_mode = mode_M > 0
        with cpu_only():
            if self.use_mode:
                self.adapter = R.dota_mixture().DOTA_mix(cfg, feat_dim, K, text.t().contiguous(), num_modes=mode_M)
            else:
                self.adapter = R.dota().DOTA(cfg, feat_dim, K, torch.full((feat_dim, K), 0.001))   # Uni_Adapter.py:329-330
        self.res_learning = res_learning and self.use_mode
        if self.res_learning:
            self.res = torch.zeros_like(text, requires_grad=True)                               # Uni_Adapter.py:346-352
            self.opt = torch.optim.Adam([self.res], lr=0.001)
        self.i = 0

    def step(self, pc, rgb):
        """pc, rgb (1,N,3) -> dict(final, clip_logits, dota_logits) (and the adapter advanced by one sample)."""
        with cpu_only():
            return self._step(pc, rgb)

    @torch.no_grad()
    def _step(self, pc, rgb):
        ua, cfg, adapter, text = self.ua, self.cfg, self.adapter, self.text
        feature = torch.cat((pc, rgb), dim=-1)
        if self.res_learning:
            clip_weights = F.normalize(text + self.res.detach(), dim=1).t()                    # :388-396
        else:
            clip_weights = text.t()
        feats, clip_logits, loss, prob_map, pred = ua.get_logits_wrapper(self.args, self.model, feature, clip_weights)
        dl = adapter.predict(feats.mean(0).unsqueeze(0).half())                                 # :410 / :416
        adapter.fit(feats, prob_map)                                                            # :411 / :417
        if self.use_mode:
            pc_aug = pc + 0.05 * torch.randn_like(pc)                                           # :420-421
            feats_aug, _, _, _, _ = ua.get_logits_wrapper(self.args, self.model, torch.cat((pc_aug, rgb), dim=-1),
                                                          clip_weights)
            feats_aug = feats_aug / feats_aug.norm(dim=-1, keepdim=True)                        # :429
            adapter.fit(feats_aug, prob_map)                                                    # :430
            adapter.update()                                                                    # :441
            if self.i > 0 and self.res_learning:                                                # :443-476
                with torch.enable_grad():
                    emb = text + self.res
                    emb = emb / emb.norm(dim=1, keepdim=True)
                    al, _ = ua.compute_text_alignment_loss(emb, adapter)
                    for _ in range(10):
                        self.opt.zero_grad()
                        al.backward()
                        self.opt.step()
                        emb = text + self.res
                        emb = emb / emb.norm(dim=1, keepdim=True)
                        al, _ = ua.compute_text_alignment_loss(emb, adapter)
            w = torch.clamp(cfg['rho'] * adapter.c.mean() / feats.size(0), max=cfg['eta'])      # :491
            d = w * dl                                                                          # :498
            ec, ed = ua.softmax_entropy(clip_logits), ua.softmax_entropy(d)                     # :508-509
            wc, wd = 1 / (ec + 1e-3), 1 / (ed + 1e-3)
            wc = wc / (wc + wd)                                                                 # :512
            wd = wd / (wc + wd)                                                                 # :513 (updated wc)
            final = wc * clip_logits + wd * d                                                   # :521
        else:
            adapter.update()                                                                    # :412
            w = torch.clamp(cfg['rho'] * adapter.c.mean() / feats.size(0), max=cfg['eta'])
            final = (clip_logits + w * dl).float()                                              # dota_mixture.py:289-293 (D2)
        self.i += 1
        return dict(final=final, clip_logits=clip_logits, dota_logits=dl.float())
