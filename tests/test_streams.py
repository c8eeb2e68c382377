"""Host-side stream ingestion: the reference's .npy corruption layout, memory-mapped, and the pinned
double-buffer prefetcher that feeds the lock-step engine."""
import numpy as np
import pytest
import torch

from uniadapter_b200.streams import (CLASS_NAMES, H5Stream, NpyCorruptionStream, PinnedPrefetcher, SyntheticStream,
                                     load_tta_dataset)


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


def test_label_layouts_mixed_corruptions_and_class_names(tmp_path):
    """data/tta_datasets.py:24-27 (mixed-corruption files), :155-163 (ScanObjectNN labels stored as [1,N] or [N,1]): every
    layout with one label per sample gives the same items; the class vocabulary follows --dataset_name like
    data/data_utils.py:11-25."""
    import types
    rng = np.random.default_rng(1)
    data = rng.standard_normal((6, 64, 3)).astype(np.float32)
    labels = rng.integers(0, 15, 6).astype(np.int64)
    for shape in ((6,), (6, 1), (1, 6)):
        root = tmp_path / f"l{len(shape)}_{shape[0]}"
        root.mkdir()
        np.save(root / "data_shear_3.npy", data)
        np.save(root / "label.npy", labels.reshape(shape))
        ds = NpyCorruptionStream(str(root), "shear", 3, dataset="scanobject_nn")
        assert [ds[i][1] for i in range(6)] == labels.tolist()
        assert ds[2][2] == CLASS_NAMES['scanobject'][labels[2]]
    root = tmp_path / "mixed"
    root.mkdir()
    np.save(root / "mixed_corruptions_5.npy", data)
    np.save(root / "mixed_corruptions_labels.npy", labels)
    ds = NpyCorruptionStream(str(root), "mixed_corruptions_5", 5, dataset="modelnet40_c")
    assert len(ds) == 6 and ds[4][1] == int(labels[4]) and ds[4][2] == CLASS_NAMES['modelnet'][labels[4]]
    args = types.SimpleNamespace(myroot=str(root), dataset_name="ModelNet", corruption="mixed_corruptions_5", severity=5,
                                 npoints=32)
    via = load_tta_dataset(args)
    assert tuple(via[0][0].shape) == (32, 3) and via[1][1] == int(labels[1])
    bad = tmp_path / "bad"
    bad.mkdir()
    np.save(bad / "data_shear_3.npy", data)
    np.save(bad / "label.npy", labels[:4])
    with pytest.raises(ValueError):
        NpyCorruptionStream(str(bad), "shear", 3)
    with pytest.raises(NotImplementedError):
        load_tta_dataset(types.SimpleNamespace(myroot=str(root), dataset_name="objaverse", corruption="x", severity=1))


def test_h5_stream_reads_like_the_reference(tmp_path, monkeypatch):
    """data/tta_datasets.py:38-95 (ModelNet_h5): file search order, float32 / int64 datasets, 1-based labels shifted.
    h5py is not installed here: a minimal stand-in module serves the two datasets from .npy files."""
    import sys
    import types
    rng = np.random.default_rng(2)
    data = rng.standard_normal((5, 32, 3)).astype(np.float64)
    labels = rng.integers(1, 41, (5, 1)).astype(np.int32)
    labels[0, 0] = 1
    np.save(tmp_path / "clean.h5.data.npy", data)
    np.save(tmp_path / "clean.h5.label.npy", labels)
    (tmp_path / "clean.h5").write_bytes(b"stand-in")

    class File:
        def __init__(self, path, mode):
            self.path = path

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def __getitem__(self, key):
            return np.load(f"{self.path}.{key}.npy")

    monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=File))
    ds = H5Stream(str(tmp_path), "clean", npoints=16)
    assert len(ds) == 5
    pc, label, name, rgb = ds[3]
    assert pc.dtype == torch.float32 and tuple(pc.shape) == (16, 3) and np.allclose(pc.numpy(), data[3, :16].astype(np.float32))
    assert label == int(labels[3, 0]) - 1 and name == CLASS_NAMES['modelnet'][label] and torch.equal(rgb, torch.ones(16, 3))
    args = types.SimpleNamespace(myroot=str(tmp_path), dataset_name="modelnet", corruption="clean", severity=5, npoints=16)
    assert isinstance(load_tta_dataset(args), H5Stream)            # no data_original.npy: the .h5 file serves 'clean'
    monkeypatch.delitem(sys.modules, "h5py")
    monkeypatch.setitem(sys.modules, "h5py", None)
    with pytest.raises(ImportError):
        H5Stream(str(tmp_path), "clean")


def test_pinned_prefetcher_propagates_a_feeder_error():
    """A dataset that fails inside the prefetch thread must fail the consuming loop (it used to leave it waiting forever)."""
    class Broken(SyntheticStream):
        def __getitem__(self, i):
            if i == 2:
                raise IndexError("label outside the class vocabulary")
            return super().__getitem__(i)

    feed = PinnedPrefetcher([Broken(5, 16, 4, seed=1)], 16)
    with pytest.raises(RuntimeError, match="stream prefetch failed"):
        list(feed)
