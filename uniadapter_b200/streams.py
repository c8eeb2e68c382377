"""Synthetic corruption streams (no datasets exist offline).

A stream is a sequence of (pc, label, class_name, rgb) samples shaped like the reference's dataset items
(data/tta_datasets.py:119-129: rgb is all ones): Gaussian clouds scaled into the unit sphere, uniform labels.
One independent generator per stream (seed + stream index), which is also the per-stream RNG convention of the
multi-GPU partitioning.
"""
from __future__ import annotations

import torch
from torch.utils.data import Dataset

CORRUPTIONS = ['uniform', 'gaussian', 'background', 'impulse', 'upsampling', 'distortion_rbf', 'distortion_rbf_inv',
               'density', 'density_inc', 'shear', 'rotation', 'cutout', 'distortion', 'occlusion', 'lidar']


def unit_sphere_clouds(n_clouds: int, npoints: int, generator: torch.Generator) -> torch.Tensor:
    x = torch.randn(n_clouds, npoints, 3, generator=generator)
    return x / x.norm(dim=-1).amax(dim=1, keepdim=True).unsqueeze(-1)


class SyntheticStream(Dataset):
    def __init__(self, length: int, npoints: int, num_classes: int, seed: int = 42, stream: int = 0,
                 colored: bool = False):
        g = torch.Generator().manual_seed(seed + stream)
        self.pc = unit_sphere_clouds(length, npoints, g)
        self.labels = torch.randint(0, num_classes, (length,), generator=g)
        self.rgb = torch.rand(length, npoints, 3, generator=g) if colored else None
        self.npoints = npoints

    def __len__(self):
        return self.pc.shape[0]

    def __getitem__(self, i):
        rgb = self.rgb[i] if self.rgb is not None else torch.ones(self.npoints, 3)
        return self.pc[i], int(self.labels[i]), f"class_{int(self.labels[i])}", rgb


def synthetic_text_features(num_classes: int, dim: int, seed: int = 0) -> torch.Tensor:
    """(K,D) unit-norm rows standing in for the CLIP text anchors (the LVIS fixture is missing upstream)."""
    g = torch.Generator().manual_seed(10_000 + seed)
    t = torch.randn(num_classes, dim, generator=g)
    return t / t.norm(dim=-1, keepdim=True)


# ----------------------------------------------------------------------------------------------------------
# real corruption files: data/tta_datasets.py:11-36,98-132
# ----------------------------------------------------------------------------------------------------------
class NpyCorruptionStream(Dataset):
    """ModelNet40-C / ScanObjectNN-C / ShapeNet-C stream from the reference's file layout
    (``data_{corruption}_{severity}.npy`` + ``label.npy``, or ``data_original.npy`` for 'clean'; data/tta_datasets.py:11-36).
    The file is memory-mapped (``mmap_mode='r'``): samples are cut out on access, nothing is loaded up front; items have
    the reference's shape (pointcloud (N,3) float32, label int, class name, rgb = ones, data/tta_datasets.py:119-129)."""

    def __init__(self, root: str, corruption: str, severity: int = 5, class_names=None, npoints: int | None = None):
        import os
        import numpy as np
        name = 'data_original.npy' if corruption == 'clean' else f'data_{corruption}_{severity}.npy'
        data_file, label_file = os.path.join(root, name), os.path.join(root, 'label.npy')
        for f in (data_file, label_file):
            if not os.path.exists(f):
                raise FileNotFoundError(f"Data file not found: {f}")
        self.data = np.load(data_file, mmap_mode='r')
        self.label = np.load(label_file, mmap_mode='r')
        self.class_names = class_names
        self.npoints = npoints

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        import numpy as np
        pc = torch.from_numpy(np.ascontiguousarray(self.data[i][: self.npoints, :3], dtype=np.float32))
        label = int(np.asarray(self.label[i]).reshape(-1)[0])
        name = self.class_names[label] if self.class_names else f"class_{label}"
        return pc, label, name, torch.ones_like(pc)


class PinnedPrefetcher:
    """Feeds ``StreamEngine.step`` from S datasets: a background thread assembles the next (S,N,3) batch into one of a few
    pinned host buffers while the GPU works on the current one, so the host->device copy of a step always starts from
    pinned memory that is already filled (the engine's copy is asynchronous on its stream)."""

    def __init__(self, datasets, npoints: int, depth: int = 2, with_rgb: bool = False):
        """``with_rgb``: every item is (xyz, labels, rgb) instead of (xyz, labels), rgb in its own pinned (S,N,3) buffer:
        for coloured streams (OpenShape); every dataset class of the reference returns rgb = ones."""
        import queue
        import threading
        self.datasets, self.N = datasets, npoints
        self.S = len(datasets)
        self.with_rgb = with_rgb
        self.length = min(len(d) for d in datasets)

        def pinned():
            t = torch.empty(self.S, npoints, 3)
            return t.pin_memory() if torch.cuda.is_available() else t
        self.buffers = [(pinned(), pinned() if with_rgb else None) for _ in range(depth + 1)]
        self.free = queue.Queue()
        for b in self.buffers:
            self.free.put(b)
        self.ready = queue.Queue(maxsize=depth)
        self.thread = threading.Thread(target=self._fill, daemon=True)
        self.thread.start()

    def _fill(self):
        for i in range(self.length):
            buf = self.free.get()
            labels = []
            for s, d in enumerate(self.datasets):
                pc, label, _, rgb = d[i]
                buf[0][s].copy_(pc[: self.N])
                if self.with_rgb:
                    buf[1][s].copy_(rgb[: self.N])
                labels.append(label)
            self.ready.put((buf, torch.tensor(labels)))
        self.ready.put(None)

    def __iter__(self):
        prev = None
        while True:
            item = self.ready.get()
            if prev is not None:
                self.free.put(prev)          # the engine has consumed it (step() synchronises on its result)
            if item is None:
                self.thread.join()           # a daemon thread still alive at interpreter shutdown aborts the process
                return
            prev = item[0]
            if self.with_rgb:
                yield item[0][0], item[1], item[0][1]
            else:
                yield item[0][0], item[1]
