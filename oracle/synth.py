"""TEST INFRASTRUCTURE — deterministic synthetic inputs shared by oracle/make_golden.py, the tests and bench.py.

numpy PCG64 streams only (pure-C ziggurat / integer paths, bit-identical on every host), with norms accumulated by a
sequential float64 cumsum so that no SIMD-width-dependent summation order can change an fp32 rounding. Golden files
store crc32(inputs) so that a host that regenerates different bits is detected instead of silently mis-compared.
"""
from __future__ import annotations

import zlib

import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.PCG64(int(seed)))


def _norm64(x):
    x64 = np.asarray(x, dtype=np.float64)
    return np.sqrt(np.cumsum(x64 * x64, axis=-1)[..., -1])


def cloud(B, N, seed):
    """Gaussian cloud scaled into the unit sphere (max point norm 1), fp32 (B,N,3). SURVEY §8d."""
    x = _rng(seed).standard_normal((B, N, 3), dtype=np.float32)
    scale = _norm64(x).max(axis=1)[:, None, None]
    return (x.astype(np.float64) / scale).astype(np.float32)


def unit_rows(n, d, seed):
    x = _rng(seed).standard_normal((n, d), dtype=np.float32)
    return (x.astype(np.float64) / _norm64(x)[:, None]).astype(np.float32)


def uniform(shape, seed):
    return _rng(seed).random(shape, dtype=np.float32)


def integers(low, high, shape, seed):
    return _rng(seed).integers(low, high, size=shape, dtype=np.int64)


def features(T, B, D, text, seed, noise=0.7, aug=0.2):
    """Unit-norm stand-ins for encoder outputs: a text row plus noise, and an 'augmented' twin. (T,B,D) each."""
    r = _rng(seed)
    K = text.shape[0]
    lab = r.integers(0, K, size=(T, B))
    x = text[lab].astype(np.float64) + noise * r.standard_normal((T, B, D), dtype=np.float32) / np.sqrt(D)
    xa = x + aug * r.standard_normal((T, B, D), dtype=np.float32) / np.sqrt(D)
    x = (x / _norm64(x)[..., None]).astype(np.float32)
    xa = (xa / _norm64(xa)[..., None]).astype(np.float32)
    return x, xa, lab


def crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return np.uint32(c)
