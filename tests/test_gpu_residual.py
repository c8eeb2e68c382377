"""GPU: residual text-feature learning kernels (csrc/residual.cu) against the reference's autograd goldens, the CPU
oracle and a plain PyTorch fp32 autograd + torch.optim.Adam run of the same loop."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import adapters as A
from oracle import cases

pytestmark = pytest.mark.gpu


def cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("name", ["align_k40_m8_d512", "align_k15_m4_d128"])
def test_alignment_loss_grad_vs_reference_autograd(name, cuda_device):
    import uniadapter_b200 as ua
    from uniadapter_b200.residual import align_loss_grad
    inp = cases.align_inputs(name)
    gold = load_golden(name, inp)
    dev = cuda_device
    loss, lm, grad, emb = align_loss_grad(cu(inp["text"], dev), cu(inp["residual"], dev), cu(gold["mu"], dev),
                                          cu(gold["var"], dev), cu(gold["pi"], dev), cases.CFG['epsilon'])
    # likelihoods are O(1e3) sums over D (fp32 ulp 1e-4..2e-4); the loss is a ratio of exp(exp(.)) of them
    np.testing.assert_allclose(lm[0].cpu().numpy(), gold["likelihood"], rtol=1e-5, atol=1e-4 * inp["D"] ** 0.5)
    np.testing.assert_allclose(loss[0].item(), gold["loss"], rtol=1e-5)
    scale = np.abs(gold["grad"]).max()
    np.testing.assert_allclose(grad[0].cpu().numpy(), gold["grad"], rtol=1e-3, atol=1e-3 * scale)
    o_loss, o_lm, o_grad, o_emb = A.align_loss_grad(inp["text"], inp["residual"], gold["mu"], gold["var"], gold["pi"],
                                                    cases.CFG['epsilon'])
    np.testing.assert_allclose(emb[0].cpu().numpy(), o_emb, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(grad[0].cpu().numpy(), o_grad, rtol=1e-3, atol=1e-3 * scale)


def test_alignment_unsupported_shape_fails_loudly(cuda_device):
    from uniadapter_b200 import _lib
    from uniadapter_b200.residual import align_loss_grad
    inp = cases.align_inputs("align_k10_m4_d64")          # D = 64: not a multiple of 128
    gold = load_golden("align_k10_m4_d64", inp)
    with pytest.raises(_lib.UaError):
        align_loss_grad(cu(inp["text"], cuda_device), cu(inp["residual"], cuda_device), cu(gold["mu"], cuda_device),
                        cu(gold["var"], cuda_device), cu(gold["pi"], cuda_device), cases.CFG['epsilon'])


def _warm_state(K, M, D, S, dev, seed):
    """S MODE-DOTA caches after a few fits of synthetic features (stacked state)."""
    from uniadapter_b200.engine import MultiStreamModeDota
    from oracle import synth
    text = cu(synth.unit_rows(K, D, seed), dev)
    cache = MultiStreamModeDota(cases.CFG, D, K, text, M, S, dev)
    gen = torch.Generator().manual_seed(seed)
    for _ in range(4):
        x = torch.nn.functional.normalize(text[torch.randint(0, K, (S,), generator=gen)]
                                          + 0.05 * torch.randn(S, D, generator=gen).to(dev), dim=-1).unsqueeze(1)
        g = torch.softmax(100.0 * x @ text.t(), -1).contiguous()
        cache.step(None, x.contiguous(), g)
    return text, cache


@pytest.mark.parametrize("shape", [(40, 8, 512, 3), (15, 8, 1280, 2), (55, 4, 256, 1), (216, 8, 512, 1), (300, 4, 512, 1),
                                   (130, 4, 256, 2)])
def test_residual_learner_vs_torch_autograd_adam(shape, cuda_device):
    """Two consecutive learn() calls (2 x 10 Adam steps, bias corrections continue) against torch autograd of the
    reference-shaped loss + torch.optim.Adam, stream by stream. K = 216 (OmniObject3D) and K = 300 take the large-K
    forms: likelihood matrix read from global memory / P in global scratch, column-chunked backward kernel."""
    from uniadapter_b200.residual import ResidualLearner, compute_text_alignment_loss
    K, M, D, S = shape
    dev = cuda_device
    text, cache = _warm_state(K, M, D, S, dev, seed=11)
    learner = ResidualLearner(text, S, M, dev)
    losses = torch.zeros(S, 10, device=dev)
    for _ in range(2):
        learner.learn(cache.mu, cache.var, cache.pi, cache.epsilon, iters=10, loss_out=losses)
    assert int(learner.adam_t[0]) == 20

    class View:       # the reference-shaped model object compute_text_alignment_loss expects
        def __init__(self, s, dtype):
            self.mu, self.var, self.pi = cache.mu[s].to(dtype), cache.var[s].to(dtype), cache.pi[s].to(dtype)
            self.eps = cache.epsilon

        def _get_var(self):
            return torch.clamp(self.var + self.eps, min=1e-8)

        def _log_likelihood(self, x, mu, var):
            diff = x.unsqueeze(1).unsqueeze(2) - mu.unsqueeze(0)
            return -0.5 * (torch.sum(torch.log(var.unsqueeze(0)), dim=-1) + torch.sum(diff ** 2 / var.unsqueeze(0), dim=-1))

    def torch_loop(s, dtype):
        t0 = text.to(dtype)
        res = torch.zeros(K, D, device=dev, dtype=dtype, requires_grad=True)
        opt = torch.optim.Adam([res], lr=1e-3)
        ls = []
        for _ in range(20):
            emb = t0 + res
            emb = emb / emb.norm(dim=1, keepdim=True)
            loss, _ = compute_text_alignment_loss(emb, View(s, dtype))
            ls.append(float(loss.detach()))
            opt.zero_grad()
            loss.backward()
            opt.step()
        return res.detach(), torch.nn.functional.normalize(t0 + res.detach(), dim=1), np.array(ls)

    # Adam moves every element by ~lr per step whatever the size of its gradient, so elements whose gradient is
    # numerically zero follow rounding noise: the yardstick is the distance of torch's OWN fp32 run from a float64 run
    # of the same loop. The CUDA path must be as close to float64 as torch fp32 is (x3), and the text rows the head sees
    # must agree to 1e-5 in cosine.
    for s in range(S):
        r64, t64, l64 = torch_loop(s, torch.float64)
        r32, t32, l32 = torch_loop(s, torch.float32)
        d_ref = (r32.double() - r64).abs()
        d_our = (learner.residual[s].double() - r64).abs()
        stats = dict(ours_mean=float(d_our.mean()), torch32_mean=float(d_ref.mean()), ours_max=float(d_our.max()),
                     torch32_max=float(d_ref.max()), moved=float(r64.abs().mean()))
        assert stats["ours_mean"] <= 3.0 * stats["torch32_mean"] + 1e-7, stats
        assert stats["ours_max"] <= 3.0 * stats["torch32_max"] + 1e-3, stats
        loss_err = np.abs(losses[s].cpu().numpy() - l64[10:]).max()
        assert loss_err <= 3.0 * np.abs(l32[10:] - l64[10:]).max() + 1e-5, (loss_err, np.abs(l32 - l64).max())
        cos = (learner.text[s].double() * t64).sum(-1)
        cos32 = (t32.double() * t64).sum(-1)          # torch's own fp32 run against float64: the yardstick again
        assert 1 - float(cos.min()) < max(1e-5, 3.0 * (1 - float(cos32.min()))), (float(cos.min()), float(cos32.min()))


def test_residual_learner_refresh_only(cuda_device):
    from uniadapter_b200.residual import ResidualLearner
    from oracle import synth
    text = cu(synth.unit_rows(12, 256, 5), cuda_device)
    learner = ResidualLearner(text, 2, 4, cuda_device)
    np.testing.assert_allclose(learner.text[1].cpu().numpy(), text.cpu().numpy(), rtol=1e-6, atol=1e-8)
    assert float(learner.residual.abs().max()) == 0.0 and int(learner.adam_t.sum()) == 0
