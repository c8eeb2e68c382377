"""fp32-accurate tensor-core GEMM (csrc/gemm_tf32x3.cu): C = epilogue(A @ W^T) with A [M,K], W [N,K] (nn.Linear /
1x1-conv weight layout), operands carried as (hi, lo) tf32 pairs."""
from __future__ import annotations

import torch

from . import _lib


def split_tf32(x: torch.Tensor):
    """x (fp32, contiguous) -> (hi, lo): hi = tf32(x) with the low 13 mantissa bits zero, lo = x - hi."""
    x = x.float().contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    rc = _lib.lib().ua_split_tf32_f32(_lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(), _lib.stream_ptr())
    _lib.check(rc, "ua_split_tf32_f32")
    return hi, lo


_ACT = {None: 0, 'none': 0, 'relu': 1, 'gelu': 2}


def layernorm_split(x, ln, pos=None, want_sum=False):
    """(hi, lo) of ``ln(x + pos)`` over the last dimension; with ``want_sum`` also returns ``x + pos``."""
    x = x.contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    s = torch.empty_like(x) if want_sum else None
    rc = _lib.lib().ua_layernorm_split_f32(_lib.ptr(x), _lib.ptr(pos.contiguous()) if pos is not None else None,
                                          _lib.ptr(ln.weight), _lib.ptr(ln.bias), float(ln.eps), rows, C, _lib.ptr(s),
                                          _lib.ptr(hi), _lib.ptr(lo), _lib.stream_ptr())
    _lib.check(rc, "ua_layernorm_split_f32")
    return ((hi, lo), s) if want_sum else (hi, lo)


def gemm_tf32x3(a, w, bias=None, group_bias=None, relu=False, out=False, out_split=False, group_max=False,
                group_max_split=False, act=None, residual=None):
    """a = (a_hi, a_lo) [M,K]; w = (w_hi, w_lo) [N,K]; epilogue + bias [N] + group_bias [M/32,N] + residual [M,N], then
    ReLU / erf-GELU (``act``). Returns a dict with the requested outputs:
    'out' [M,N] fp32, 'out_split' (hi, lo), 'gmax' [M/32,N], 'gmax_split' (hi, lo)."""
    a_hi, a_lo = a
    w_hi, w_lo = w
    M, K = a_hi.shape
    N = w_hi.shape[0]
    dev = a_hi.device
    res = {}
    o = torch.empty((M, N), dtype=torch.float32, device=dev) if out else None
    oh = torch.empty((M, N), dtype=torch.float32, device=dev) if out_split else None
    ol = torch.empty((M, N), dtype=torch.float32, device=dev) if out_split else None
    gm = torch.empty((M // 32, N), dtype=torch.float32, device=dev) if (group_max or group_max_split) else None
    gh = torch.empty_like(gm) if group_max_split else None
    gl = torch.empty_like(gm) if group_max_split else None
    rc = _lib.lib().ua_gemm_tf32x3_f32(
        _lib.ptr(a_hi), _lib.ptr(a_lo), a_hi.stride(0), _lib.ptr(w_hi), _lib.ptr(w_lo), w_hi.stride(0), M, N, K,
        _lib.ptr(bias), _lib.ptr(group_bias), _lib.ptr(residual), 1 if relu else _ACT[act], _lib.ptr(o), _lib.ptr(oh),
        _lib.ptr(ol), N, _lib.ptr(gm),
        _lib.ptr(gh), _lib.ptr(gl), _lib.stream_ptr())
    _lib.check(rc, "ua_gemm_tf32x3_f32")
    if out:
        res['out'] = o
    if out_split:
        res['out_split'] = (oh, ol)
    if gm is not None:
        res['gmax'] = gm
    if group_max_split:
        res['gmax_split'] = (gh, gl)
    return res


# ----------------------------------------------------------------------------------------------------------
# modules built on the GEMM: the group encoder (mini-PointNet) and nn.Linear
# ----------------------------------------------------------------------------------------------------------
def _fold_bn(conv_w, conv_b, bn):
    """Conv1d(k=1) followed by BatchNorm1d in eval mode -> one affine map (W', b')."""
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = conv_w.squeeze(-1) * s[:, None]
    b = (conv_b - bn.running_mean) * s + bn.bias
    return w.contiguous(), b.contiguous()


class GroupEncoderPlan:
    """Inference plan of a ``MiniPointNet`` (eval mode) on the tensor-core GEMM (models/ulip/pointbert/dvae.py:201-215,
    models/point_encoder.py:145-159):

        h1 = relu(bn(conv1(x)))                 point-wise kernel (C = 3 / 6 input channels) -> (hi, lo)
        f  = conv2(h1);  g = max_points(f)      GEMM K=128, epilogue: (hi, lo) store + max over the 32-row groups
        h3 = relu(bn(conv3([g, f])))            = relu(f . W3f'^T + (g . W3g'^T + b3'))  — the concatenation with the
                                                broadcast global feature becomes a per-group bias (one small GEMM)
        out = max_points(conv4(h3))             GEMM K=512, epilogue: max over the 32-row groups only (conv4's output
                                                is never written)
    BatchNorm is folded into the convolutions; weights are split into (hi, lo) once."""

    def __init__(self, module):
        with torch.no_grad():
            c1, bn1, _, c2 = module.first_conv
            c3, bn3, _, c4 = module.second_conv
            self.C = c1.in_channels
            self.E = c4.out_channels
            self.w1, self.b1 = _fold_bn(c1.weight, c1.bias, bn1)                      # (128, C)
            self.w2 = split_tf32(c2.weight.squeeze(-1))                               # (256, 128)
            self.b2 = c2.bias.detach().clone().contiguous()
            w3, self.b3 = _fold_bn(c3.weight, c3.bias, bn3)                           # (512, 512): [global | local]
            self.w3g = split_tf32(w3[:, :256])
            self.w3f = split_tf32(w3[:, 256:])
            self.w4 = split_tf32(c4.weight.squeeze(-1))                               # (E, 512)
            self.b4 = c4.bias.detach().clone().contiguous()

    @torch.no_grad()
    def __call__(self, point_groups: torch.Tensor) -> torch.Tensor:
        bs, g, n, c = point_groups.shape
        if n % 32 or c != self.C:
            raise _lib.UaError(f"group encoder: group size {n} must be a multiple of 32 and channels {c} == {self.C}")
        M, r = bs * g * n, n // 32
        x = point_groups.reshape(M, c).contiguous()
        dev = x.device
        h1 = (torch.empty((M, 128), device=dev), torch.empty((M, 128), device=dev))
        rc = _lib.lib().ua_pointwise_linear_split_f32(_lib.ptr(x), _lib.ptr(self.w1), _lib.ptr(self.b1), 1, M, c, 128,
                                                       _lib.ptr(h1[0]), _lib.ptr(h1[1]), _lib.stream_ptr())
        _lib.check(rc, "ua_pointwise_linear_split_f32")
        res = gemm_tf32x3(h1, self.w2, bias=self.b2, out_split=True, group_max=(r > 1), group_max_split=(r == 1))
        f = res['out_split']
        if r == 1:
            gpair = res['gmax_split']
        else:            # groups of 64 points span two 32-row halves
            gpair = split_tf32(res['gmax'].view(-1, r, 256).amax(1))
        gbias = gemm_tf32x3(gpair, self.w3g, bias=self.b3, out=True)['out']                  # (bs*g, 512)
        if r > 1:
            gbias = gbias.repeat_interleave(r, 0).contiguous()
        h3 = gemm_tf32x3(f, self.w3f, group_bias=gbias, relu=True, out_split=True)['out_split']
        out = gemm_tf32x3(h3, self.w4, bias=self.b4, group_max=True)['gmax']                 # (M/32, E)
        if r > 1:
            out = out.view(-1, r, self.E).amax(1)
        return out.reshape(bs, g, self.E)


class LinearPlan:
    """nn.Linear on the tensor-core GEMM: y = x @ W^T + b with W split once; x is split per call."""

    def __init__(self, linear):
        with torch.no_grad():
            self.w = split_tf32(linear.weight)
            self.b = linear.bias.detach().clone().contiguous() if linear.bias is not None else None
            self.N, self.K = linear.weight.shape

    @staticmethod
    def supported(linear) -> bool:
        n, k = linear.weight.shape
        return n % 128 == 0 and k % 32 == 0

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, relu=False) -> torch.Tensor:
        lead = x.shape[:-1]
        a = split_tf32(x.reshape(-1, self.K))
        return gemm_tf32x3(a, self.w, bias=self.b, relu=relu, out=True)['out'].reshape(*lead, self.N)


def attention_tf32x3(qkv: torch.Tensor, B: int, N: int, H: int):
    """qkv [B*N, 3*H*64] fp32 (columns ordered q|k|v, head, d) -> (hi, lo) of softmax(q k^T / 8) v, [B*N, H*64]
    (csrc/attention.cu: tcgen05, P kept in tensor memory)."""
    dev = qkv.device
    qkv = qkv.contiguous()
    npad = int(_lib.lib().ua_attn_padded_tokens(N))
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    q_hi, q_lo, k_hi, k_lo = (e(B * H, N, 64) for _ in range(4))
    vt_hi, vt_lo = e(B * H, 64, npad), e(B * H, 64, npad)
    out_hi, out_lo = e(B * N, H * 64), e(B * N, H * 64)
    rc = _lib.lib().ua_attn_prepare_f32(_lib.ptr(qkv), B, N, H, _lib.ptr(q_hi), _lib.ptr(q_lo), _lib.ptr(k_hi),
                                       _lib.ptr(k_lo), _lib.ptr(vt_hi), _lib.ptr(vt_lo), _lib.stream_ptr())
    _lib.check(rc, "ua_attn_prepare_f32")
    rc = _lib.lib().ua_attention_f32(_lib.ptr(q_hi), _lib.ptr(q_lo), _lib.ptr(k_hi), _lib.ptr(k_lo), _lib.ptr(vt_hi),
                                    _lib.ptr(vt_lo), B, N, H, _lib.ptr(out_hi), _lib.ptr(out_lo), _lib.stream_ptr())
    _lib.check(rc, "ua_attention_f32")
    return out_hi, out_lo


class BlockPlan:
    """A pre-LN transformer block with every dense layer on the tensor-core GEMM and the element-wise work fused around
    it: (x + pos, LayerNorm, split) in one kernel, the skip connections in the proj / fc2 epilogues, GELU + split in the
    fc1 epilogue. Attention runs on csrc/attention.cu when the head dimension is 64 (torch SDPA otherwise)."""

    def __init__(self, block, tc_attention=True):
        self.block = block
        self.tc_attention = tc_attention
        self.qkv, self.proj = LinearPlan(block.attn.qkv), LinearPlan(block.attn.proj)
        self.fc1, self.fc2 = LinearPlan(block.mlp.fc1), LinearPlan(block.mlp.fc2)

    @staticmethod
    def supported(block) -> bool:
        dims_ok = block.norm1.normalized_shape[0] % 128 == 0 and block.norm1.normalized_shape[0] <= 1024
        return dims_ok and all(LinearPlan.supported(l) for l in (block.attn.qkv, block.attn.proj, block.mlp.fc1, block.mlp.fc2))

    @torch.no_grad()
    def __call__(self, x, pos=None):
        import torch.nn.functional as F
        blk = self.block
        B, N, C = x.shape
        H = blk.attn.heads
        h, xs = layernorm_split(x, blk.norm1, pos, want_sum=True) if pos is not None else (layernorm_split(x, blk.norm1), x)
        flat = lambda pair: (pair[0].view(B * N, -1), pair[1].view(B * N, -1))
        qkv = gemm_tf32x3(flat(h), self.qkv.w, bias=self.qkv.b, out=True)['out']
        if C // H == 64 and self.tc_attention:
            a = attention_tf32x3(qkv, B, N, H)               # tcgen05 attention, already the (hi, lo) pair proj needs
        else:
            q, k, v = qkv.view(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
            a = split_tf32(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, C))
        x1 = gemm_tf32x3(a, self.proj.w, bias=self.proj.b, residual=xs.view(B * N, C), out=True)['out']
        h2 = layernorm_split(x1, blk.norm2)
        g = gemm_tf32x3(h2, self.fc1.w, bias=self.fc1.b, act='gelu', out_split=True)['out_split']
        x2 = gemm_tf32x3(g, self.fc2.w, bias=self.fc2.b, residual=x1, out=True)['out']
        return x2.view(B, N, C)
