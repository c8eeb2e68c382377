"""Point-cloud encoders around the sm_100a tokenizer.

The north star keeps the transformer blocks in PyTorch (cuBLAS / SDPA): they are timed, not optimised. What changes is
the front end — every encoder below tokenizes through ``uniadapter_b200.tokenizer`` (fps.cu / group.cu) instead of the
G-iteration torch FPS loop, the materialised (G,N) distance matrix, topk / sort and the index kernels.

* ``UlipPointBert``   — ULIP-2 PointBERT (reference: models/ulip/pointbert/point_encoder.py:103-192,
  dvae.py:185-215, ulip_model.py:8-18; hyper-parameters PointTransformer_8192point.yaml:15-25).
  Parameters are created in the reference's order with the same initialisers, so ``torch.manual_seed(s)`` yields
  the same random-init weights as the reference module (checked by oracle/make_golden.py).
* ``Uni3DEncoder``    — Uni3D point encoder (models/point_encoder.py:129-223, models/uni3d.py:9-19). The reference
  takes its blocks from timm's EVA02 (not installed, and not part of the parity target): a seeded pre-LN
  transformer of the same width / depth / heads stands in.
* ``OpenShapePPAT``   — OpenShape PointPatchTransformer, scaling 4 (models/openshape/ppta.py:85-148,181-186,
  pointnet_util.py:165-210): set abstraction = FPS + ball query + (9→64→64→256 conv, max).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import tokenizer as tok


# ----------------------------------------------------------------------------------------------------------
# shared blocks
# ----------------------------------------------------------------------------------------------------------
class MiniPointNet(nn.Module):
    """Per-group encoder: conv(C→128) BN ReLU conv(128→256) | max | concat | conv(512→512) BN ReLU conv(512→E) | max."""

    def __init__(self, in_channels: int, encoder_channel: int):
        super().__init__()
        self.encoder_channel = encoder_channel
        self.first_conv = nn.Sequential(nn.Conv1d(in_channels, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                                        nn.Conv1d(128, 256, 1))
        self.second_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                                         nn.Conv1d(512, encoder_channel, 1))

    def forward(self, point_groups: torch.Tensor) -> torch.Tensor:
        plan = getattr(self, '_tc_plan', None)
        if plan is not None and not self.training and point_groups.is_cuda and not torch.is_grad_enabled():
            return plan(point_groups)                       # tcgen05 path (gemm.GroupEncoderPlan)
        bs, g, n, c = point_groups.shape
        f = self.first_conv(point_groups.reshape(bs * g, n, c).transpose(2, 1))
        f = torch.cat([f.max(dim=2, keepdim=True)[0].expand(-1, -1, n), f], dim=1)
        f = self.second_conv(f).max(dim=2)[0]
        return f.reshape(bs, g, self.encoder_channel)


def _linear(mod: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """nn.Linear through its tensor-core plan when ``use_tensor_cores`` attached one (inference on CUDA), else torch."""
    plan = getattr(mod, '_tc_plan', None)
    if plan is not None and x.is_cuda and not torch.is_grad_enabled():
        return plan(x)
    return mod(x)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return _linear(self.fc2, F.gelu(_linear(self.fc1, x)))


class _Attention(nn.Module):
    def __init__(self, dim, heads, qkv_bias=False):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = _linear(self.qkv, x).reshape(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        x = F.scaled_dot_product_attention(q, k, v)
        return _linear(self.proj, x.transpose(1, 2).reshape(B, N, C))


class _Block(nn.Module):
    """Pre-LN block; sub-modules are created in the reference's order (norm1, norm2, mlp, attn)."""

    def __init__(self, dim, heads, mlp_ratio=4.0, qkv_bias=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.attn = _Attention(dim, heads, qkv_bias)

    def forward(self, x, pos=None):
        """``pos`` (the position embedding PointBERT re-adds before every block) is added first: blk(x + pos)."""
        plan = getattr(self, '_tc_plan', None)
        if plan is not None and x.is_cuda and not torch.is_grad_enabled():
            return plan(x, pos)                              # fused tensor-core path (gemm.BlockPlan)
        if pos is not None:
            x = x + pos
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


# ----------------------------------------------------------------------------------------------------------
# ULIP-2 PointBERT
# ----------------------------------------------------------------------------------------------------------
class _PointBertTrunk(nn.Module):
    def __init__(self, trans_dim, depth, heads, group_size, num_group, encoder_dim):
        super().__init__()
        self.group_divider = tok.Group(num_group, group_size, random_start=True)
        self.encoder = MiniPointNet(3, encoder_dim)
        self.reduce_dim = nn.Linear(encoder_dim, trans_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, trans_dim))
        self.cls_pos = nn.Parameter(torch.randn(1, 1, trans_dim))
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, trans_dim))
        self.blocks = nn.ModuleList([_Block(trans_dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(trans_dim)

    def forward(self, pts):
        neighborhood, center = self.group_divider(pts)
        tokens = _linear(self.reduce_dim, self.encoder(neighborhood))
        B = tokens.size(0)
        x = torch.cat((self.cls_token.expand(B, -1, -1), tokens), dim=1)
        pos = torch.cat((self.cls_pos.expand(B, -1, -1), self.pos_embed(center)), dim=1)
        for blk in self.blocks:
            x = blk(x, pos)           # the reference re-adds the position embedding before every block: blk(x + pos)
        x = self.norm(x)
        return torch.cat([x[:, 0], x[:, 1:].max(1)[0]], dim=-1)


class UlipPointBert(nn.Module):
    """ULIP-2 point branch: PointBERT trunk (768-d concat feature) @ pc_projection (768, 512)."""

    def __init__(self, pc_feat_dim=768, depth=12, heads=6, group_size=32, num_group=512, encoder_dim=256, embed_dim=512):
        super().__init__()
        self.point_encoder = _PointBertTrunk(pc_feat_dim // 2, depth, heads, group_size, num_group, encoder_dim)
        # the reference allocates this with torch.empty and fills it from a checkpoint; random init here
        self.pc_projection = nn.Parameter(torch.empty(pc_feat_dim, embed_dim))
        nn.init.normal_(self.pc_projection, std=pc_feat_dim ** -0.5)

    def forward(self, pc):
        return self.point_encoder(pc) @ self.pc_projection


# ----------------------------------------------------------------------------------------------------------
# Uni3D
# ----------------------------------------------------------------------------------------------------------
class Uni3DEncoder(nn.Module):
    """Uni3D-L geometry: 512 groups x 64 neighbours, encoder 512 -> trans 1024 (24 blocks, 16 heads) -> embed 1024."""

    def __init__(self, pc_feat_dim=1024, embed_dim=1024, group_size=64, num_group=512, pc_encoder_dim=512, depth=24,
                 heads=16, mlp_ratio=8.0 / 3.0):
        super().__init__()
        self.group_divider = tok.Group(num_group, group_size, random_start=False)
        self.encoder = MiniPointNet(6, pc_encoder_dim)
        self.encoder2trans = nn.Linear(pc_encoder_dim, pc_feat_dim)
        self.trans2embed = nn.Linear(pc_feat_dim, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, pc_feat_dim))
        self.cls_pos = nn.Parameter(torch.randn(1, 1, pc_feat_dim))
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, pc_feat_dim))
        self.blocks = nn.ModuleList([_Block(pc_feat_dim, heads, mlp_ratio, qkv_bias=True) for _ in range(depth)])
        self.norm = nn.LayerNorm(pc_feat_dim)
        self.fc_norm = nn.LayerNorm(pc_feat_dim)

    def encode_pc(self, pc):
        xyz = pc[:, :, :3].contiguous()
        color = pc[:, :, 3:].contiguous()
        return self.forward(xyz, color)

    def front(self, pts, colors):
        """Tokenizer + group encoder + position embedding: the (B, 1+G, C) sequence entering the transformer blocks
        (models/point_encoder.py:192-208)."""
        _, center, features = self.group_divider(pts, colors)
        tokens = _linear(self.encoder2trans, self.encoder(features))
        B = tokens.size(0)
        x = torch.cat((self.cls_token.expand(B, -1, -1), tokens), dim=1)
        return x + torch.cat((self.cls_pos.expand(B, -1, -1), self.pos_embed(center)), dim=1)

    def forward(self, pts, colors):
        x = self.front(pts, colors)
        for blk in self.blocks:
            x = blk(x)
        return self.trans2embed(self.fc_norm(self.norm(x[:, 0, :])))


# ----------------------------------------------------------------------------------------------------------
# OpenShape PointPatchTransformer (scaling 4)
# ----------------------------------------------------------------------------------------------------------
class _SetAbstraction(nn.Module):
    def __init__(self, npoint, radius, nsample, in_channel, mlp):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out))
            last = out
        self.next_start_idx = None
        self.device_rng = False

    def forward(self, xyz, points):
        """xyz (B,3,N), points (B,D,N) -> new_xyz (B,3,S), features (B,C,S)."""
        xyz_t = xyz.permute(0, 2, 1).contiguous()
        pts_t = points.permute(0, 2, 1).contiguous() if points is not None else None
        start, self.next_start_idx = self.next_start_idx, None
        if start is None and self.device_rng:
            start = torch.randint(0, xyz_t.shape[1], (xyz_t.shape[0],), dtype=torch.long, device=xyz_t.device)
        new_xyz, new_points = tok.sample_and_group(self.npoint, self.radius, self.nsample, xyz_t, pts_t, start_idx=start)
        h = new_points.permute(0, 3, 2, 1)           # (B, C+D, nsample, S)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            h = F.relu(bn(conv(h)))
        return new_xyz.permute(0, 2, 1), h.max(dim=2)[0]


class _PPATLayer(nn.Module):
    def __init__(self, dim, heads, dim_head, mlp_dim):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.norm_a = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Linear(inner, dim)
        self.norm_f = nn.LayerNorm(dim)
        self.ff1 = nn.Linear(dim, mlp_dim)
        self.ff2 = nn.Linear(mlp_dim, dim)

    def forward(self, x):
        B, N, _ = x.shape
        q, k, v = self.to_qkv(self.norm_a(x)).reshape(B, N, 3, self.heads, -1).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, -1)
        x = self.to_out(a) + x
        return self.ff2(F.gelu(self.ff1(self.norm_f(x)))) + x


class OpenShapePPAT(nn.Module):
    def __init__(self, dim=512, depth=12, heads=8, mlp_dim=1536, sa_dim=256, patches=384, prad=0.2, nsamp=64, in_dim=6,
                 dim_head=64, out_channel=1280):
        super().__init__()
        self.sa = _SetAbstraction(patches, prad, nsamp, in_dim + 3, [64, 64, sa_dim])
        self.lift_conv = nn.Conv1d(sa_dim + 3, dim, 1)
        self.lift_norm = nn.LayerNorm(dim)
        self.cls_token = nn.Parameter(torch.randn(dim))
        self.layers = nn.ModuleList([_PPATLayer(dim, heads, dim_head, mlp_dim) for _ in range(depth)])
        self.proj = nn.Linear(dim, out_channel)

    def forward(self, xyz, features):
        """xyz (B,N,3), features (B,N,6) -> (B, out_channel)."""
        centroids, feature = self.sa(xyz.transpose(-1, -2).contiguous(), features.transpose(-1, -2).contiguous())
        x = self.lift_norm(self.lift_conv(torch.cat([centroids, feature], dim=1)).permute(0, 2, 1))
        x = torch.cat([self.cls_token.expand(x.size(0), 1, -1), x], dim=1)
        for layer in self.layers:
            x = layer(x)
        return self.proj(x[:, 0])


def build_encoder(vlm3d: str, seed: int = 0, device='cuda', small: bool = False, tensor_cores: bool = True) -> nn.Module:
    """Random-init encoder of the named family in eval mode (no checkpoints exist offline).
    ``small`` shrinks the transformer depth for smoke tests. ``tensor_cores`` (CUDA devices) routes the group encoder
    and the supported Linear layers through the tcgen05 3xTF32 GEMM (``use_tensor_cores``)."""
    torch.manual_seed(seed)
    if vlm3d == 'ulip':
        m = UlipPointBert(depth=2 if small else 12)
    elif vlm3d == 'uni3d':
        m = Uni3DEncoder(depth=2 if small else 24)
    elif vlm3d == 'openshape':
        m = OpenShapePPAT(depth=2 if small else 12)
    else:
        raise ValueError(f"unknown vlm3d {vlm3d!r}")
    m = m.to(device).float().eval()
    if tensor_cores and torch.device(device).type == 'cuda':
        use_tensor_cores(m, True)
    return m


def use_tensor_cores(model: nn.Module, flag: bool = True, linears: bool = True) -> nn.Module:
    """Attach (or drop) the tcgen05 3xTF32 plans (gemm.py): the mini-PointNet group encoder and, with ``linears``, the
    nn.Linear layers whose shapes the GEMM supports (N % 128 == 0, K % 32 == 0). Inference only; weights are folded
    and split when this is called, so call it again after loading other weights."""
    from .gemm import BlockPlan, GroupEncoderPlan, LinearPlan
    if flag and not getattr(model, '_tc_reload_hook', None):
        # the plans snapshot (fold + split) the weights: rebuild them whenever other weights are loaded
        def _rebuild(module, incompatible_keys):      # (a post hook must return None)
            use_tensor_cores(module, True, linears)

        model._tc_reload_hook = model.register_load_state_dict_post_hook(_rebuild)
    for mod in model.modules():
        if isinstance(mod, MiniPointNet):
            mod._tc_plan = GroupEncoderPlan(mod) if flag else None
        elif isinstance(mod, _Block) and linears:
            mod._tc_plan = BlockPlan(mod) if (flag and BlockPlan.supported(mod)) else None
        elif isinstance(mod, nn.Linear) and linears:
            mod._tc_plan = LinearPlan(mod) if (flag and LinearPlan.supported(mod)) else None
    return model


def set_device_rng(model: nn.Module, flag: bool = True) -> None:
    """Draw FPS start indices on the device generator (graph-capturable) instead of the CPU one (reference order)."""
    for mod in model.modules():
        if hasattr(mod, 'device_rng'):
            mod.device_rng = flag
