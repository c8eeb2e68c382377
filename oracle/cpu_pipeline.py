"""TEST INFRASTRUCTURE — the reference's per-sample test-time step on the CPU, assembled from the oracle pieces.

Used as (a) the checker of the end-to-end parity tests and of __graft_entry__.smoke(), and (b) the CPU baseline /
``--impl reference`` arm of bench.py (timed on the GPU box's host cores). Never on the product path.

Tokenizer: oracle/tokenizer_oracle.c; head / MODE-DOTA / DOTA / fusion: oracle/adapters.py (numpy); the encoder's
PyTorch layers (mini-PointNet, transformer blocks — not part of the parity target) run on the CPU from
the same module definitions and weights as the GPU run; residual text learning restates Uni_Adapter.py:191-270,443-476
in CPU autograd with the reference's (K,K,M,D) broadcast formulation.
"""
from __future__ import annotations

import numpy as np
import torch

from . import adapters as A
from . import tokenizer as T


class OracleGroup(torch.nn.Module):
    """CPU stand-in for the group divider (FPS + kNN + centre subtraction) backed by the C oracle."""

    def __init__(self, num_group, group_size, random_start, threads=1, pointnet2=False):
        super().__init__()
        self.num_group, self.group_size, self.random_start, self.threads = num_group, group_size, random_start, threads
        self.pointnet2 = pointnet2          # Uni3D: the published pointnet2_ops FPS (oracle_fps_pointnet2)
        self.next_start_idx = None

    def forward(self, xyz, color=None):
        B, N, _ = xyz.shape
        start = None
        if self.next_start_idx is not None:
            start, self.next_start_idx = self.next_start_idx.cpu().numpy(), None
        elif self.random_start:
            start = torch.randint(0, N, (B,), dtype=torch.long).numpy()     # misc.py:52
        g = T.group_knn(xyz.numpy(), self.num_group, self.group_size, rgb=None if color is None else color.numpy(),
                        start_idx=start, threads=self.threads, sort_by_index=True,
                        pointnet2=self.pointnet2 and start is None)
        if color is None:
            return torch.from_numpy(g["neigh"]), torch.from_numpy(g["center"])
        return torch.from_numpy(g["neigh"]), torch.from_numpy(g["center"]), torch.from_numpy(g["feat"])


class OracleSetAbstractionGroup:
    """sample_and_group on the CPU (OpenShape front end)."""

    @staticmethod
    def run(npoint, radius, nsample, xyz, points, start, threads=1):
        g = T.sample_and_group(xyz.numpy(), npoint, radius, nsample, None if points is None else points.numpy(),
                               None if start is None else start.numpy(), threads)
        return torch.from_numpy(g["new_xyz"]), torch.from_numpy(g["new_points"])


def cpu_encoder_like(gpu_or_fresh_encoder, threads=1):
    """A CPU copy of a ``uniadapter_b200.encoders`` module whose tokenizer is the oracle."""
    import copy
    enc = copy.deepcopy(gpu_or_fresh_encoder).cpu().float().eval()
    for name, mod in list(enc.named_modules()):
        if mod.__class__.__name__ == "Group":
            parent = enc
            parts = name.split(".")
            for p in parts[:-1]:
                parent = getattr(parent, p)
            setattr(parent, parts[-1], OracleGroup(mod.num_group, mod.group_size, mod.random_start, threads,
                                                       getattr(mod, 'pointnet2', False)))
        if mod.__class__.__name__ == "_SetAbstraction":
            sa = mod

            def fwd(xyz, points, _sa=sa):
                xyz_t = xyz.permute(0, 2, 1).contiguous()
                pts_t = points.permute(0, 2, 1).contiguous() if points is not None else None
                start, _sa.next_start_idx = _sa.next_start_idx, None
                if start is None:
                    start = torch.randint(0, xyz_t.shape[1], (xyz_t.shape[0],), dtype=torch.long)
                new_xyz, new_points = OracleSetAbstractionGroup.run(_sa.npoint, _sa.radius, _sa.nsample, xyz_t, pts_t,
                                                                     start, threads)
                h = new_points.permute(0, 3, 2, 1)
                for conv, bn in zip(_sa.mlp_convs, _sa.mlp_bns):
                    h = torch.relu(bn(conv(h)))
                return new_xyz.permute(0, 2, 1), h.max(dim=2)[0]

            sa.forward = fwd
    return enc


def encode_cpu(encoder, vlm3d, pc, rgb):
    with torch.no_grad():
        if vlm3d == 'uni3d':
            return encoder.encode_pc(torch.cat((pc, rgb), dim=-1))
        if vlm3d == 'ulip':
            return encoder(pc)
        return encoder(pc, torch.cat((pc, rgb), dim=-1))


def alignment_loss_cpu(emb, model: A.ModeDota):
    """Uni_Adapter.py:191-270 on CPU autograd: (K,K,M,D) broadcast likelihood, double-exp contrastive loss."""
    mu = torch.from_numpy(model.mu)
    v = torch.from_numpy(model.reg_var())
    diff = emb.unsqueeze(1).unsqueeze(2) - mu.unsqueeze(0)
    ll = -0.5 * (torch.log(v).sum(-1).unsqueeze(0) + (diff ** 2 / v.unsqueeze(0)).sum(-1))
    lm = torch.logsumexp(torch.log(torch.from_numpy(model.pi) + 1e-10).unsqueeze(0) + ll, dim=2)
    e = torch.exp(torch.exp(lm / lm.max()))
    d = torch.diag(e)
    return -(d / e.sum(1)).mean() - (d / e.sum(0)).mean(), lm


class CpuStream:
    """One corruption stream adapted on the CPU exactly as Uni_Adapter.py:368-579 does (batch 1)."""

    def __init__(self, encoder_cpu, vlm3d, text, cfg, adapter='mode_dota', M=8, res_learning=True):
        self.enc, self.vlm3d, self.cfg = encoder_cpu, vlm3d, cfg
        self.text0 = np.asarray(text, dtype=np.float32)
        K, D = self.text0.shape
        self.kind, self.res_learning = adapter, res_learning and adapter == 'mode_dota'
        if adapter == 'mode_dota':
            self.model = A.ModeDota(cfg, D, K, self.text0.T, M)
        else:
            self.model = A.Dota(cfg, D, K, np.full((D, K), 0.001, dtype=np.float32))
        self.i = 0
        if self.res_learning:
            self.res = torch.zeros(K, D, requires_grad=True)
            self.opt = torch.optim.Adam([self.res], lr=0.001)

    def current_text(self):
        if not self.res_learning:
            return self.text0
        t = torch.from_numpy(self.text0) + self.res.detach()
        return torch.nn.functional.normalize(t, dim=1).numpy()

    def step(self, pc, rgb, noise=None, start=None, start_aug=None):
        """pc, rgb (1,N,3) CPU tensors -> dict(final, pred, clip_logits, dota_logits)."""
        cfg = self.cfg
        text = self.current_text()
        self._inject(start)
        h = A.head(encode_cpu(self.enc, self.vlm3d, pc, rgb).numpy(), text)
        x = h["xnorm"]
        xp = x.mean(axis=0, keepdims=True, dtype=np.float32).astype(np.float16)
        if self.kind == 'dota':
            dl = self.model.predict(xp)
            self.model.fit(x, h["prob"])
            self.model.update()
            final, _ = A.fuse_dota(h["logits"], dl, self.model.c, cfg['rho'], cfg['eta'], x.shape[0])
        else:
            dl = self.model.predict(xp.astype(np.float32))
            self.model.fit(x, h["prob"])
            if noise is None:
                noise = torch.randn_like(pc)
            self._inject(start_aug)
            ha = A.head(encode_cpu(self.enc, self.vlm3d, pc + 0.05 * noise, rgb).numpy(), text)
            self.model.fit(ha["xnorm"], h["prob"])
            if self.i > 0 and self.res_learning:
                t0 = torch.from_numpy(self.text0)
                for it in range(11):
                    emb = t0 + self.res
                    emb = emb / emb.norm(dim=1, keepdim=True)
                    loss, _ = alignment_loss_cpu(emb, self.model)
                    if it == 10:
                        break
                    self.opt.zero_grad()
                    loss.backward()
                    self.opt.step()
            final, _ = A.fuse_mode_dota(h["logits"], dl, self.model.c, cfg['rho'], cfg['eta'], x.shape[0])
        self.i += 1
        return dict(final=final, pred=final.argmax(axis=1), clip_logits=h["logits"], dota_logits=np.asarray(dl))

    def _inject(self, start):
        if start is None:
            return
        for mod in self.enc.modules():
            if hasattr(mod, "next_start_idx"):
                mod.next_start_idx = torch.as_tensor(start)
