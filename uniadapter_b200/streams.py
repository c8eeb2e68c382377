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
# real corruption files: data/tta_datasets.py, data/data_utils.py
# ----------------------------------------------------------------------------------------------------------
CLASS_NAMES = {    # the label vocabularies of the reference's dataset classes (data/tta_datasets.py:69-77,151-154,250-261)
    'modelnet': ["airplane", "bathtub", "bed", "bench", "bookshelf", "bottle", "bowl", "car", "chair", "cone", "cup", "curtain",
                 "desk", "door", "dresser", "flower_pot", "glass_box", "guitar", "keyboard", "lamp", "laptop", "mantel",
                 "monitor", "night_stand", "person", "piano", "plant", "radio", "range_hood", "sink", "sofa", "stairs",
                 "stool", "table", "tent", "toilet", "tv_stand", "vase", "wardrobe", "xbox"],
    'scanobject': ["bag", "bin", "box", "cabinet", "chair", "desk", "display", "door", "shelf", "table", "bed", "pillow", "sink",
                   "sofa", "toilet"],
    'shapenet': ["airplane", "bag", "basket", "bathtub", "bed", "bench", "bottle", "bowl", "bus", "cabinet", "can", "camera",
                 "cap", "car", "chair", "clock", "dishwasher", "monitor", "table", "telephone", "tin_can", "tower", "train",
                 "keyboard", "earphone", "faucet", "file", "guitar", "helmet", "jar", "knife", "lamp", "laptop", "speaker",
                 "mailbox", "microphone", "microwave", "motorcycle", "mug", "piano", "pillow", "pistol", "pot", "printer",
                 "remote_control", "rifle", "rocket", "skateboard", "sofa", "stove", "vessel", "washer", "cellphone",
                 "birdhouse", "bookshelf"],
}


def _dataset_family(name: str | None) -> str | None:
    name = (name or '').lower()
    for fam in ('modelnet', 'scanobject', 'shapenet'):          # the same substring test as data/data_utils.py:11-25
        if fam in name:
            return fam
    return None


def corruption_files(root: str, corruption: str, severity: int):
    """File names of the reference's ``load_data`` (data/tta_datasets.py:11-36): ``data_{corruption}_{severity}.npy`` +
    ``label.npy``; ``data_original.npy`` for 'clean'; ``{corruption}.npy`` + ``mixed_corruptions_labels.npy`` for the
    mixed-corruption streams."""
    import os
    if 'mixed_corruptions' in corruption:
        return os.path.join(root, f'{corruption}.npy'), os.path.join(root, 'mixed_corruptions_labels.npy')
    name = 'data_original.npy' if corruption == 'clean' else f'data_{corruption}_{severity}.npy'
    return os.path.join(root, name), os.path.join(root, 'label.npy')


class NpyCorruptionStream(Dataset):
    """ModelNet40-C / ScanObjectNN-C / ShapeNet-C stream from the reference's file layout (``corruption_files``).
    The data file is memory-mapped (``mmap_mode='r'``): samples are cut out on access, nothing is loaded up front; items
    have the reference's shape (pointcloud (N,3) float32, label int, class name, rgb = ones, data/tta_datasets.py:119-129).

    Labels are read layout-agnostically like the reference does (``ModelNet40C`` takes ``label[i]`` and unwraps a
    1-element array, ``ScanObjectNN_C`` tries ``label[0][i]`` first because its labels come as ``[1,N]`` or ``[N,1]``,
    data/tta_datasets.py:155-163): any array with one label per sample -- ``(N,)``, ``(N,1)`` or ``(1,N)`` -- is flattened.
    ``dataset`` ('modelnet' | 'scanobject' | 'shapenet', or any name containing one of them, as ``--dataset_name``) selects
    the class-name vocabulary; unknown names give ``class_{label}``."""

    def __init__(self, root: str, corruption: str, severity: int = 5, class_names=None, npoints: int | None = None,
                 dataset: str | None = None, debug: bool = False):
        import os
        import numpy as np
        data_file, label_file = corruption_files(root, corruption, severity)
        for f in (data_file, label_file):
            if not os.path.exists(f):
                raise FileNotFoundError(f"{'Data' if f == data_file else 'Label'} file not found: {f}")
        self.data = np.load(data_file, mmap_mode='r')
        label = np.asarray(np.load(label_file, allow_pickle=True))
        n = self.data.shape[0]
        if label.size != n:
            raise ValueError(f"{label_file}: {label.size} labels for {n} samples (shape {label.shape})")
        self.label = label.reshape(-1).astype(np.int64)          # (N,), (N,1) and (1,N) all mean one label per sample
        if debug:                                                # data/tta_datasets.py:106-108
            self.data, self.label = self.data[:5], self.label[:5]
        self.family = _dataset_family(dataset)
        self.class_names = class_names if class_names is not None else CLASS_NAMES.get(self.family)
        self.npoints = npoints

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        import numpy as np
        pc = torch.from_numpy(np.array(self.data[i][: self.npoints, :3], dtype=np.float32))
        label = int(self.label[i])
        name = self.class_names[label] if self.class_names else f"class_{label}"
        return pc, label, name, torch.ones_like(pc)


class H5Stream(Dataset):
    """Clean ModelNet40 from an HDF5 file as the reference's ``ModelNet_h5`` reads it (data/tta_datasets.py:38-95): the
    first of ``modelnet40_test.h5``, ``clean.h5``, ``{corruption}.h5`` found under ``root``; datasets ``data`` (float32) and
    ``label`` (int64); 1-based labels are shifted to 0-based. Needs ``h5py`` (absent from this image: ImportError says so)."""

    def __init__(self, root: str, corruption: str = 'clean', npoints: int | None = None, class_names=None):
        import os
        import numpy as np
        try:
            import h5py
        except ImportError as exc:
            raise ImportError("H5Stream needs h5py to read .h5 streams (not installed in this image); convert the file to the "
                              ".npy layout of NpyCorruptionStream or install h5py") from exc
        names = ['modelnet40_test.h5', 'clean.h5', f'{corruption}.h5']
        path = next((os.path.join(root, n) for n in names if os.path.exists(os.path.join(root, n))), None)
        if path is None:
            raise FileNotFoundError(f"Could not find H5 file in {root}. Checked: {names}")
        with h5py.File(path, 'r') as f:
            self.data = np.asarray(f['data'][:], dtype=np.float32)
            self.label = np.asarray(f['label'][:], dtype=np.int64).reshape(-1)
        if self.label.size and int(self.label.min()) == 1:
            self.label = self.label - 1
        self.class_names = class_names if class_names is not None else CLASS_NAMES['modelnet']
        self.npoints = npoints

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        pc = torch.from_numpy(self.data[i][: self.npoints, :3].copy())
        label = int(self.label[i])
        return pc, label, self.class_names[label], torch.ones_like(pc)


def load_tta_dataset(args):
    """data/data_utils.py:5-26: the dataset of ``args.dataset_name`` / ``args.corruption`` / ``args.severity`` under
    ``args.myroot`` (ModelNet40-C, ScanObjectNN-C and ShapeNet-C share the .npy layout; the reference keeps its .h5 class
    for clean ModelNet40 but routes 'clean' to the .npy loader too -- the .h5 file is read when the .npy one is absent)."""
    import os
    fam = _dataset_family(args.dataset_name)
    if fam is None:
        raise NotImplementedError(f'Dataset {args.dataset_name} is not implemented')
    data_file, _ = corruption_files(args.myroot, args.corruption, args.severity)
    if fam == 'modelnet' and args.corruption == 'clean' and not os.path.exists(data_file):
        return H5Stream(args.myroot, args.corruption, npoints=getattr(args, 'npoints', None))
    return NpyCorruptionStream(args.myroot, args.corruption, args.severity, npoints=getattr(args, 'npoints', None),
                               dataset=fam, debug=bool(getattr(args, 'debug', False)))


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
        try:
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
        except BaseException as exc:      # noqa: BLE001  a dead feeder must fail the consumer, not leave it waiting
            self.ready.put(exc)

    def __iter__(self):
        prev = None
        while True:
            item = self.ready.get()
            if isinstance(item, BaseException):
                self.thread.join()
                raise RuntimeError(f"stream prefetch failed: {type(item).__name__}: {item}") from item
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
