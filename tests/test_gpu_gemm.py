"""GPU: the tcgen05 3xTF32 GEMM, the tensor-core group encoder and Linear plans against plain PyTorch references
(float64 for the GEMM itself, the fp32 torch modules for the plans)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(128, 128, 32), (100, 256, 64), (1000, 384, 384), (2048, 512, 256), (4097, 128, 1536)])
def test_gemm_tf32x3_vs_float64(shape, cuda_device):
    """fp32-class accuracy: the error against float64 stays within a small multiple of an fp32 matmul's own error."""
    from uniadapter_b200.gemm import gemm_tf32x3, split_tf32
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda_device)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device)
    b = torch.randn(N, generator=g).to(cuda_device)
    ap, wp = split_tf32(a), split_tf32(w)
    assert torch.equal(ap[0] + ap[1], a) and int((ap[0].view(torch.int32) & 0x1fff).abs().max()) == 0
    out = gemm_tf32x3(ap, wp, bias=b, out=True)['out']
    ref = a.double() @ w.double().t() + b.double()
    err = float((out.double() - ref).abs().max())
    err32 = float(((a @ w.t() + b).double() - ref).abs().max())
    assert err <= 1e-5 * float(ref.abs().max()) and err <= 12 * err32 + 1e-6, (err, err32)


def test_gemm_epilogues(cuda_device):
    from uniadapter_b200.gemm import gemm_tf32x3, split_tf32
    M, N, K = 4096, 256, 128
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g).to(cuda_device)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device)
    b = torch.randn(N, generator=g).to(cuda_device)
    gb = torch.randn(M // 32, N, generator=g).to(cuda_device)
    res = gemm_tf32x3(split_tf32(a), split_tf32(w), bias=b, group_bias=gb, relu=True, out=True, out_split=True,
                      group_max_split=True)
    ref = torch.relu(a.double() @ w.double().t() + b.double() + gb.double().repeat_interleave(32, 0)).float()
    np.testing.assert_allclose(res['out'].cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=2e-5)
    hi, lo = res['out_split']
    assert torch.equal(hi + lo, res['out']) and int((hi.view(torch.int32) & 0x1fff).abs().max()) == 0
    assert torch.equal(res['gmax'], res['out'].view(M // 32, 32, N).amax(1))
    gh, gl = res['gmax_split']
    assert torch.equal(gh + gl, res['gmax'])


def test_gemm_rejects_unsupported_shapes(cuda_device):
    from uniadapter_b200 import _lib
    from uniadapter_b200.gemm import gemm_tf32x3, split_tf32
    a = split_tf32(torch.randn(64, 48, device=cuda_device))
    w = split_tf32(torch.randn(100, 48, device=cuda_device))
    with pytest.raises(_lib.UaError):
        gemm_tf32x3(a, w, out=True)


@pytest.mark.parametrize("cfg", [(3, 256, 32, 2, 50), (6, 512, 64, 1, 24)])
def test_group_encoder_plan_vs_torch_module(cfg, cuda_device):
    """MiniPointNet on the tensor cores (BN folded, concat as per-group bias, fused group max) vs the fp32 torch module
    evaluated in float64."""
    from uniadapter_b200.encoders import MiniPointNet
    from uniadapter_b200.gemm import GroupEncoderPlan
    C, E, n, bs, g = cfg
    torch.manual_seed(5)
    mod = MiniPointNet(C, E).to(cuda_device).eval()
    with torch.no_grad():
        for bn in (mod.first_conv[1], mod.second_conv[1]):        # non-trivial running statistics
            bn.running_mean.normal_(0, 0.1), bn.running_var.uniform_(0.5, 1.5), bn.weight.normal_(1, 0.1), bn.bias.normal_(0, 0.1)
    x = torch.randn(bs, g, n, C, device=cuda_device) * 0.3
    plan = GroupEncoderPlan(mod)
    with torch.no_grad():
        out = plan(x)
        ref = mod.double()(x.double()).float()
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-5)


def test_encoder_with_tensor_cores_matches_torch_encoder(cuda_device):
    """The whole ULIP encoder with tensor-core plans (group encoder + Linear layers) vs the same encoder in torch fp32
    with TF32 disabled: features within fp32 round-off, far tighter than the cuDNN TF32 convolutions torch uses."""
    from uniadapter_b200.encoders import UlipPointBert, use_tensor_cores
    from uniadapter_b200.streams import unit_sphere_clouds
    torch.manual_seed(0)
    enc = UlipPointBert(depth=2).to(cuda_device).eval()
    pc = unit_sphere_clouds(3, 1024, torch.Generator().manual_seed(2)).to(cuda_device)
    start = torch.zeros(3, dtype=torch.long, device=cuda_device)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            enc.point_encoder.group_divider.next_start_idx = start
            ref = enc(pc)
            use_tensor_cores(enc, True)
            enc.point_encoder.group_divider.next_start_idx = start
            out = enc(pc)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    scale = float(ref.abs().max())
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=2e-5 * scale)


def test_layernorm_split_and_fused_epilogues(cuda_device):
    """(x + pos) -> LayerNorm -> (hi, lo); GELU / residual epilogues of the GEMM; against torch in float64."""
    from uniadapter_b200.gemm import gemm_tf32x3, layernorm_split, split_tf32
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 77, 384, generator=g).to(dev)
    pos = torch.randn(5, 77, 384, generator=g).to(dev)
    ln = torch.nn.LayerNorm(384).to(dev)
    with torch.no_grad():
        ln.weight.normal_(1, 0.2), ln.bias.normal_(0, 0.2)
        (hi, lo), s = layernorm_split(x, ln, pos, want_sum=True)
        ref = ln.double()((x + pos).double()).float()
    assert torch.equal(s, x + pos)
    np.testing.assert_allclose((hi + lo).cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert int((hi.view(torch.int32) & 0x1fff).abs().max()) == 0
    M, N, K = 385, 256, 384
    a = (hi + lo).view(M, K)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    r = torch.randn(M, N, generator=g).to(dev)
    pre = a.double() @ w.double().t() + b.double()
    out = gemm_tf32x3((hi.view(M, K), lo.view(M, K)), split_tf32(w), bias=b, act='gelu', out=True, out_split=True)
    np.testing.assert_allclose(out['out'].cpu().numpy(), torch.nn.functional.gelu(pre).float().cpu().numpy(), rtol=1e-5, atol=5e-5)
    assert torch.equal(out['out_split'][0] + out['out_split'][1], out['out'])
    out = gemm_tf32x3((hi.view(M, K), lo.view(M, K)), split_tf32(w), bias=b, residual=r, out=True)['out']
    np.testing.assert_allclose(out.cpu().numpy(), (pre + r.double()).float().cpu().numpy(), rtol=1e-5, atol=5e-5)


def test_block_plan_vs_torch_block(cuda_device):
    from uniadapter_b200.encoders import _Block
    from uniadapter_b200.gemm import BlockPlan
    dev = cuda_device
    torch.manual_seed(4)
    blk = _Block(384, 6).to(dev).eval()
    x = torch.randn(3, 513, 384, device=dev)
    pos = torch.randn(3, 513, 384, device=dev) * 0.1
    with torch.no_grad():
        out = BlockPlan(blk)(x, pos)
        ref = blk.double()(x.double(), pos.double()).float()
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("shape", [(64, 1024, 55), (64, 1024, 1156), (200, 512, 216)])
def test_tensor_core_head_vs_simt_head_and_golden(shape, cuda_device):
    """Batched zero-shot head on the tcgen05 GEMM (cfg 5 shapes) against the SIMT head kernel; the golden b64 case too."""
    import uniadapter_b200 as ua
    from uniadapter_b200.head import HeadPlan
    from oracle import synth
    B, D, K = shape
    text = torch.from_numpy(synth.unit_rows(K, D, 9)).to(cuda_device)
    x = torch.randn(B, D, generator=torch.Generator().manual_seed(B)).to(cuda_device)
    xn, logits, ent, prob, arg = HeadPlan(text)(x)
    xn0, logits0, ent0, prob0, arg0 = ua.zero_shot_head(x, text, tensor_cores=False)      # the SIMT kernels
    xn1, logits1, ent1, prob1, arg1 = ua.zero_shot_head(x, text)                          # default: B >= 64 -> tcgen05
    assert torch.equal(logits1, logits) and torch.equal(prob1, prob) and torch.equal(arg1, arg)
    assert torch.equal(xn, xn0)
    np.testing.assert_allclose(logits.cpu().numpy(), logits0.cpu().numpy(), rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(prob.cpu().numpy(), prob0.cpu().numpy(), rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(ent.cpu().numpy(), ent0.cpu().numpy(), rtol=2e-4, atol=1e-6)
    assert torch.equal(arg, arg0)


def test_tensor_core_head_vs_reference_golden(cuda_device):
    from conftest import load_golden
    from oracle import cases
    from uniadapter_b200.head import HeadPlan
    name = "head_b64_d1024_k55"
    inp = cases.head_inputs(name)
    gold = load_golden(name, inp)
    dev = cuda_device
    xn, logits, ent, prob, arg = HeadPlan(torch.from_numpy(inp["text"]).to(dev))(torch.from_numpy(inp["x"]).to(dev))
    np.testing.assert_allclose(xn.cpu().numpy(), gold["xnorm"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(logits.cpu().numpy(), gold["logits"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(prob.cpu().numpy(), gold["prob"], rtol=2e-4, atol=1e-7)
    np.testing.assert_array_equal(arg.cpu().numpy(), gold["pred"])


@pytest.mark.parametrize("shape", [(1, 128, 1), (2, 130, 2), (3, 513, 6), (2, 700, 4), (2, 2, 1), (1, 257, 3)])
def test_tensor_core_attention_vs_sdpa_float64(shape, cuda_device):
    """csrc/attention.cu (3xTF32, P in tensor memory, base-2 online softmax) vs torch SDPA evaluated in float64."""
    import torch.nn.functional as F
    from uniadapter_b200.gemm import attention_tf32x3
    B, N, H = shape
    C = H * 64
    qkv = torch.randn(B * N, 3 * C, generator=torch.Generator().manual_seed(N)).to(cuda_device)
    hi, lo = attention_tf32x3(qkv, B, N, H)
    assert int((hi.view(torch.int32) & 0x1fff).abs().max()) == 0
    q, k, v = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double()).transpose(1, 2).reshape(B * N, C)
    np.testing.assert_allclose((hi + lo).cpu().numpy(), ref.float().cpu().numpy(), rtol=2e-5, atol=1e-5)


def test_tensor_core_plans_follow_load_state_dict(cuda_device):
    """The tcgen05 plans snapshot (fold + split) the weights when they are attached; loading other weights must rebuild them
    (a later load_state_dict used to be ignored silently)."""
    from uniadapter_b200.encoders import UlipPointBert, use_tensor_cores
    from uniadapter_b200.streams import unit_sphere_clouds
    dev = cuda_device
    torch.manual_seed(1)
    a = use_tensor_cores(UlipPointBert(depth=1).to(dev).eval(), True)
    torch.manual_seed(2)
    donor = UlipPointBert(depth=1).to(dev).eval()
    pc = unit_sphere_clouds(2, 1024, torch.Generator().manual_seed(3)).to(dev)
    start = torch.zeros(2, dtype=torch.long, device=dev)

    def run(m):
        m.point_encoder.group_divider.next_start_idx = start
        with torch.no_grad():
            return m(pc)
    before = run(a)
    a.load_state_dict(donor.state_dict())
    after = run(a)
    want = run(use_tensor_cores(donor, True))
    assert not torch.allclose(before, after, rtol=1e-3, atol=1e-3)
    assert torch.equal(after, want)
