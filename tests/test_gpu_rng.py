"""GPU: the per-stream counter-based generator of the lock-step engine (csrc/rng.cu)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def draw(seeds, step, S, per, n_range, dev):
    from uniadapter_b200 import _lib
    seeds_t = torch.tensor(seeds, dtype=torch.int64, device=dev)
    step_t = torch.tensor([step], dtype=torch.int64, device=dev)
    done = torch.zeros(1, dtype=torch.int32, device=dev)
    noise = torch.full((S, per), float('nan'), device=dev)
    start = torch.full((2, S), -1, dtype=torch.int64, device=dev)
    rc = _lib.lib().ua_stream_rng_f32(_lib.ptr(seeds_t), _lib.ptr(step_t), S, per, _lib.ptr(noise), _lib.ptr(start), n_range,
                                      _lib.ptr(done), _lib.stream_ptr())
    _lib.check(rc, "ua_stream_rng_f32")
    torch.cuda.synchronize()
    return noise.cpu(), start.cpu(), int(step_t.item()), int(done.item())


def test_stream_rng_moments_ranges_and_counter(cuda_device):
    S, per = 5, 3 * 10000 + 1          # ragged: not a multiple of four
    noise, start, step, done = draw([11, 12, 13, 14, 15], 7, S, per, 1024, cuda_device)
    assert torch.isfinite(noise).all()
    assert step == 8 and done == 0      # the kernel advances its own step counter (graph replays keep counting)
    assert ((start >= 0) & (start < 1024)).all()
    x = noise.double()
    assert abs(float(x.mean())) < 0.02 and abs(float(x.var()) - 1.0) < 0.02
    assert abs(float((x ** 3).mean())) < 0.05 and abs(float((x ** 4).mean()) - 3.0) < 0.15
    # streams are decorrelated
    c = np.corrcoef(noise.numpy())
    assert np.abs(c - np.eye(S)).max() < 0.03


def test_stream_rng_depends_only_on_own_seed_and_step(cuda_device):
    per = 3 * 1024
    a_noise, a_start, _, _ = draw([42, 43, 44], 3, 3, per, 1024, cuda_device)
    b_noise, b_start, _, _ = draw([44, 42], 3, 2, per, 1024, cuda_device)        # other company, other slot
    assert torch.equal(a_noise[0], b_noise[1]) and torch.equal(a_noise[2], b_noise[0])
    assert torch.equal(a_start[:, 0], b_start[:, 1]) and torch.equal(a_start[:, 2], b_start[:, 0])
    c_noise, _, _, _ = draw([42], 4, 1, per, 1024, cuda_device)                   # next step: new draws
    assert not torch.equal(a_noise[0], c_noise[0])
