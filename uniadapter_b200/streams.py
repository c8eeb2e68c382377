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
