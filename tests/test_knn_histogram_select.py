"""CPU restatement (numpy) of the first-tile selection of csrc/group.cu: a 256-bin histogram over a monotone map of the
distance brackets the k-th smallest, every point of a bin <= b* is a candidate, the surplus (largest (distance bits,
index) pairs) is dropped. Must equal the oracle's kNN (expanded squared distance, ties to the lower index) exactly -
the bins only pre-partition. The kernel itself is tested bit-exactly on the GPU (tests/test_gpu_tokenizer.py)."""
import numpy as np
import pytest

from oracle import synth
from oracle import tokenizer as T


def ordered_bits(d):
    u = d.astype(np.float32).view(np.uint32).astype(np.uint64)
    neg = (u & 0x80000000) != 0
    return np.where(neg, (~u) & 0xffffffff, u | 0x80000000)


def histogram_select(d, k):
    """d: (N<=1024,) fp32 distances of one centre -> sorted indices of the k smallest (d, index) pairs."""
    d = d.astype(np.float32)
    dmax = np.float32(max(d.max(), 0.0))
    delta = np.float32(dmax * np.float32(0.015625))
    v = (np.maximum(d, np.float32(0)) + delta).astype(np.float32)
    e = v.view(np.uint32).astype(np.int64) - np.array(delta, dtype=np.float32).view(np.uint32).astype(np.int64)
    bins = np.minimum(e >> 18, 255)
    hist = np.bincount(bins, minlength=256)
    cum = np.cumsum(hist)
    bstar = int(np.searchsorted(cum, k))            # first bin whose cumulative count reaches k
    cand = np.nonzero(bins <= bstar)[0]              # ascending index order
    keys = (ordered_bits(d[cand]) << np.uint64(32)) | cand.astype(np.uint64)
    drop = len(cand) - k
    if drop > 0:
        cand = cand[np.argsort(keys, kind="stable")[:k]]
    return np.sort(cand), len(cand) + max(drop, 0) - k


@pytest.mark.parametrize("N,G,k,dups", [(1024, 64, 64, False), (1024, 48, 32, False), (257, 40, 9, False), (300, 32, 16, True)])
def test_histogram_selection_equals_oracle_knn(N, G, k, dups):
    xyz = synth.cloud(1, N, 500 + N + k)
    if dups:
        xyz[:, 100:200] = xyz[:, 0:100]
        xyz[:, 200:300] = xyz[:, 0:100]
    centers = np.ascontiguousarray(xyz[:, :: max(1, N // G)][:, :G])
    want = np.sort(T.knn(xyz, centers, k, threads=2), axis=-1)[0]
    d = T.sqdist(centers[0], xyz[0])                 # (G, N): the reference's expanded form in fp32
    surplus = []
    for g in range(centers.shape[1]):
        got, extra = histogram_select(d[g], k)
        np.testing.assert_array_equal(got, want[g])
        surplus.append(extra)
    assert max(surplus) <= 40                        # the bins bracket the k-th smallest tightly (kernel falls back above 12)
