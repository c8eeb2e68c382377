"""CPU restatement (numpy, fp32) of the algorithm of csrc/spdinv.cu - block Gauss-Jordan inversion of an SPD matrix without
pivoting, 16-wide pivot blocks, in-place bookkeeping by whole micro-block rows / columns, 16x16 pivot inverse with the
next pivot's reciprocal computed one step ahead - against the float64 inverse and LAPACK's fp32 LU (the reference's
torch.inverse, dota.py:68). Documents the arithmetic the kernel implements; the kernel itself is tested on the GPU
(tests/test_gpu_adapters.py::test_dota_update_inverse_kernel)."""
import numpy as np
import pytest
import torch

F = np.float32


def invert16_lookahead(a):
    """Unpivoted scalar Gauss-Jordan on a 16x16 block; p_{t+1} = 1 / fma(-a[t+1][t], a[t][t+1] * p_t, a[t+1][t+1])."""
    a = a.astype(F).copy()
    n = a.shape[0]
    p = F(1.0) / a[0, 0]
    for t in range(n):
        old = a.copy()
        pn = F(0)
        if t + 1 < n:
            pn = F(1.0) / F(old[t + 1, t + 1] - F(old[t + 1, t] * F(old[t, t + 1] * p)))
        rp = (old[t, :] * p).astype(F)
        a = (old - np.outer(old[:, t], rp)).astype(F)
        a[t, :] = rp
        a[:, t] = (-old[:, t] * p).astype(F)
        a[t, t] = p
        p = pn
    return a


def block_gauss_jordan(A, nb=16):
    A = A.astype(F).copy()
    D = A.shape[0]
    for k0 in range(0, D, nb):
        ks = slice(k0, k0 + nb)
        P = invert16_lookahead(A[ks, ks])
        R = (P @ A[ks, :]).astype(F)                 # R = P * A[k, :]; the pivot columns of R are P itself
        R[:, ks] = P
        C = A[:, ks].copy()
        C[ks, :] = 0                                 # pivot rows take R below, not the update
        base = A.copy()
        base[:, ks] = 0                              # pivot columns become -A[:, k] * P
        new = (base - C @ R).astype(F)
        new[ks, :] = R
        A = new
    return A


def spd_like_dota(D, steps, seed):
    rng = np.random.default_rng(seed)
    A = np.eye(D) * 1e-4
    for _ in range(steps):
        d = rng.standard_normal(D) / np.sqrt(D) * 0.7
        A = 0.98 * A + 0.02 * np.outer(d, d)
    return (0.9999 * A + 1e-4 * np.eye(D)).astype(F)


@pytest.mark.parametrize("D,steps", [(16, 3), (64, 50), (256, 200)])
def test_block_gauss_jordan_matches_lapack_accuracy(D, steps):
    A = spd_like_dota(D, steps, 7 + D)
    ref = np.linalg.inv(A.astype(np.float64))
    scale = np.abs(ref).max()
    got = block_gauss_jordan(A)
    lu = torch.inverse(torch.from_numpy(A)).numpy()
    err, lu_err = np.abs(got - ref).max() / scale, np.abs(lu - ref).max() / scale
    assert err <= max(6 * lu_err, 2e-6), (err, lu_err)
    assert np.abs(A.astype(np.float64) @ got.astype(np.float64) - np.eye(D)).max() < 1e-5 * np.linalg.cond(A.astype(np.float64))
