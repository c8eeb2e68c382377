"""Host-side stream ingestion: the reference's .npy corruption layout, memory-mapped, and the pinned
double-buffer prefetcher that feeds the lock-step engine."""
import numpy as np
import torch

from uniadapter_b200.streams import NpyCorruptionStream, PinnedPrefetcher, SyntheticStream


def test_npy_corruption_stream_reads_the_reference_layout(tmp_path):
    rng = np.random.default_rng(0)
    data = rng.standard_normal((7, 1024, 3)).astype(np.float32)
    labels = rng.integers(0, 40, (7, 1)).astype(np.int64)
    np.save(tmp_path / "data_gaussian_5.npy", data)
    np.save(tmp_path / "data_original.npy", data * 2)
    np.save(tmp_path / "label.npy", labels)
    ds = NpyCorruptionStream(str(tmp_path), "gaussian", 5)
    assert len(ds) == 7
    pc, label, name, rgb = ds[3]
    assert pc.dtype == torch.float32 and tuple(pc.shape) == (1024, 3) and np.array_equal(pc.numpy(), data[3])
    assert label == int(labels[3, 0]) and name == f"class_{label}" and torch.equal(rgb, torch.ones(1024, 3))
    clean = NpyCorruptionStream(str(tmp_path), "clean", npoints=512)
    assert tuple(clean[0][0].shape) == (512, 3) and np.array_equal(clean[0][0].numpy(), 2 * data[0, :512])
    try:
        NpyCorruptionStream(str(tmp_path), "lidar", 5)
        raise AssertionError("missing file must raise")
    except FileNotFoundError:
        pass


def test_pinned_prefetcher_delivers_every_step_in_order():
    streams = [SyntheticStream(9, 64, 10, seed=1, stream=s) for s in range(3)]
    steps = list(PinnedPrefetcher(streams, 64))
    assert len(steps) == 9
    for i, (batch, labels) in enumerate(steps[-2:], start=7):      # buffers are recycled: check the freshest ones
        for s in range(3):
            assert int(labels[s]) == streams[s][i][1]
    first = PinnedPrefetcher(streams, 64)
    batch, labels = next(iter(first))
    for s in range(3):
        assert torch.equal(batch[s], streams[s][0][0])


def test_pinned_prefetcher_with_rgb():
    streams = [SyntheticStream(5, 48, 7, seed=2, stream=s, colored=True) for s in range(2)]
    n = 0
    for i, (xyz, labels, rgb) in enumerate(PinnedPrefetcher(streams, 48, with_rgb=True)):
        for s in range(2):
            assert torch.equal(xyz[s], streams[s][i][0]) and torch.equal(rgb[s], streams[s][i][3])
            assert int(labels[s]) == streams[s][i][1]
        n += 1
    assert n == 5
