"""Point tokenizer: farthest-point sampling, kNN / ball-query grouping, centre normalisation.

Host-side mirror of the reference's tokenizer call signatures; every function launches the sm_100a kernels of
libua_b200.so (fps.cu, group.cu) through the C ABI. Reference seams:

* ``pointnet2_utils.furthest_point_sample / gather_operation`` and ``fps(data, number)``  models/point_encoder.py:7-14
* ``fps(xyz, npoint)`` (random start, returns points)                                     models/ulip/pointbert/misc.py:40-60
* ``knn_point(nsample, xyz, new_xyz)``                                                    models/ulip/pointbert/dvae.py:116-127
* ``Group(num_group, group_size).forward``        models/point_encoder.py:93-127, models/ulip/pointbert/dvae.py:152-181
* ``farthest_point_sample / query_ball_point / sample_and_group``                         models/openshape/pointnet_util.py:64-146
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib

_FPS_MAX_REG_POINTS = 16384
UA_FPS_POINTNET2 = 2      # include/ua_b200.h


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ----------------------------------------------------------------------------------------------------------
# raw ops
# ----------------------------------------------------------------------------------------------------------
def fps_sample(xyz: torch.Tensor, npoint: int, start_idx: torch.Tensor | None = None, *,
               skip_small_norm: bool = False, idx_dtype: torch.dtype = torch.int64,
               want_idx: bool = True, want_centers: bool = True, pointnet2: bool = False):
    """FPS over every cloud of ``xyz`` (B,N,3). Returns (idx (B,G) | None, centers (B,G,3) | None).

    ``start_idx`` (B,) int64 gives the first sample of each cloud (None -> 0, the pointnet2_ops convention).
    ``pointnet2``: the published pointnet2_ops kernel's own arithmetic (FMA-contracted distances, near-origin points
    never take part, ties resolved like upstream's left-biased reduction tree) instead of the torch FPS arithmetic of misc.py / pointnet_util.py.
    """
    xyz = _f32c(xyz)
    B, N, ch = xyz.shape
    if ch != 3:
        raise ValueError(f"fps_sample expects (B,N,3), got {tuple(xyz.shape)}")
    if idx_dtype not in (torch.int32, torch.int64):
        raise ValueError("idx_dtype must be int32 or int64")
    idx = torch.empty((B, npoint), dtype=idx_dtype, device=xyz.device) if want_idx else None
    centers = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device) if want_centers else None
    scratch = torch.empty((B, N), dtype=torch.float32, device=xyz.device) if N > _FPS_MAX_REG_POINTS else None
    if start_idx is not None:
        start_idx = start_idx.to(device=xyz.device, dtype=torch.int64).contiguous()
        if start_idx.numel() != B:
            raise ValueError("start_idx must have one entry per cloud")
    rc = _lib.lib().ua_fps_f32(_lib.ptr(xyz), B, N, int(npoint), _lib.ptr(start_idx),
                               UA_FPS_POINTNET2 if pointnet2 else int(bool(skip_small_norm)),
                               _lib.ptr(idx), int(idx_dtype == torch.int64), _lib.ptr(centers), _lib.ptr(scratch),
                               _lib.stream_ptr())
    _lib.check(rc, "ua_fps_f32")
    return idx, centers


def knn_group(xyz: torch.Tensor, centers: torch.Tensor, k: int, rgb: torch.Tensor | None = None, *,
              want_idx: bool = False, want_neigh: bool = True, want_feat: bool | None = None,
              idx_dtype: torch.dtype = torch.int64):
    """k nearest points of every centre + gather + centre subtraction (+ colour concat).

    Returns (idx (B,G,k) | None, neigh (B,G,k,3) | None, feat (B,G,k,6) | None); neighbours in ascending
    point-index order (the reference's topk(sorted=False) order is unspecified).
    """
    xyz = _f32c(xyz)
    centers = _f32c(centers)
    B, N, _ = xyz.shape
    G = centers.shape[1]
    if want_feat is None:
        want_feat = rgb is not None
    if rgb is not None:
        rgb = _f32c(rgb)
    dev = xyz.device
    idx = torch.empty((B, G, k), dtype=idx_dtype, device=dev) if want_idx else None
    neigh = torch.empty((B, G, k, 3), dtype=torch.float32, device=dev) if want_neigh else None
    feat = torch.empty((B, G, k, 6), dtype=torch.float32, device=dev) if want_feat else None
    rc = _lib.lib().ua_knn_group_f32(_lib.ptr(xyz), _lib.ptr(rgb), _lib.ptr(centers), B, N, G, int(k), _lib.ptr(idx),
                                     int(idx_dtype == torch.int64), _lib.ptr(neigh), _lib.ptr(feat),
                                     _lib.stream_ptr())
    _lib.check(rc, "ua_knn_group_f32")
    return idx, neigh, feat


def ball_group(xyz: torch.Tensor, centers: torch.Tensor, radius: float, nsample: int,
               points: torch.Tensor | None = None, *, want_idx: bool = False, want_points: bool = True,
               idx_dtype: torch.dtype = torch.int64):
    """Ball query (first nsample in index order, padded with the first hit) + gather + centre subtraction.

    Returns (idx (B,S,nsample) | None, new_points (B,S,nsample,3+C) | None).
    """
    xyz = _f32c(xyz)
    centers = _f32c(centers)
    B, N, _ = xyz.shape
    S = centers.shape[1]
    Cf = 0
    if points is not None:
        points = _f32c(points)
        Cf = points.shape[-1]
    dev = xyz.device
    idx = torch.empty((B, S, nsample), dtype=idx_dtype, device=dev) if want_idx else None
    out = torch.empty((B, S, nsample, 3 + Cf), dtype=torch.float32, device=dev) if want_points else None
    # the reference compares fp32 distances against the Python double radius**2, i.e. against its fp32 rounding
    r2 = float(torch.tensor(float(radius) ** 2, dtype=torch.float64).to(torch.float32))
    rc = _lib.lib().ua_ball_group_f32(_lib.ptr(xyz), _lib.ptr(points), Cf, _lib.ptr(centers), B, N, S, r2,
                                      int(nsample), _lib.ptr(idx), int(idx_dtype == torch.int64), _lib.ptr(out),
                                      _lib.stream_ptr())
    _lib.check(rc, "ua_ball_group_f32")
    return idx, out


# ----------------------------------------------------------------------------------------------------------
# reference-named entry points
# ----------------------------------------------------------------------------------------------------------
def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """pointnet2_utils.furthest_point_sample: (B,N,3) -> (B,npoint) int32, first sample = point 0, in the published
    kernel's arithmetic and tie order (sampling_gpu.cu of pointnet2_ops_lib 3.0.0)."""
    idx, _ = fps_sample(xyz, npoint, None, idx_dtype=torch.int32, want_centers=False, pointnet2=True)
    return idx


def gather_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """pointnet2_utils.gather_operation: features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint)."""
    features = _f32c(features)
    idx = idx.to(torch.int32).contiguous()
    B, Cc, N = features.shape
    G = idx.shape[1]
    out = torch.empty((B, Cc, G), dtype=torch.float32, device=features.device)
    rc = _lib.lib().ua_gather_points_f32(_lib.ptr(features), _lib.ptr(idx), B, Cc, N, G, _lib.ptr(out),
                                         _lib.stream_ptr())
    _lib.check(rc, "ua_gather_points_f32")
    return out


def fps_uni3d(data: torch.Tensor, number: int) -> torch.Tensor:
    """models/point_encoder.py:7-14 fps(data, number): pointnet2_ops FPS from point 0, returns the sampled points (B,G,3)."""
    _, centers = fps_sample(data, number, None, want_idx=False, pointnet2=True)
    return centers


def fps(xyz: torch.Tensor, npoint: int, start_idx: torch.Tensor | None = None) -> torch.Tensor:
    """models/ulip/pointbert/misc.py:40-60 fps(xyz, npoint): random first sample, returns points (B,npoint,3).

    The start index is drawn with the same torch call as the reference (``torch.randint(0, N, (B,))`` on the
    global CPU generator) unless the caller supplies it.
    """
    B, N, _ = xyz.shape
    if start_idx is None:
        start_idx = torch.randint(0, N, (B,), dtype=torch.long)
    _, centers = fps_sample(xyz, npoint, start_idx, want_idx=False)
    return centers


def farthest_point_sample(xyz: torch.Tensor, npoint: int, start_idx: torch.Tensor | None = None) -> torch.Tensor:
    """models/openshape/pointnet_util.py:64-86: random first sample, returns indices (B,npoint) int64."""
    B, N, _ = xyz.shape
    if start_idx is None:
        start_idx = torch.randint(0, N, (B,), dtype=torch.long)
    idx, _ = fps_sample(xyz, npoint, start_idx, want_centers=False)
    return idx


def knn_point(nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """knn_point(nsample, xyz, new_xyz) -> (B,S,nsample) int64 (ascending index; the reference's order is unspecified)."""
    idx, _, _ = knn_group(xyz, new_xyz, nsample, want_idx=True, want_neigh=False, want_feat=False)
    return idx


def query_ball_point(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """models/openshape/pointnet_util.py:89-110 -> (B,S,nsample) int64."""
    idx, _ = ball_group(xyz, new_xyz, radius, nsample, want_idx=True, want_points=False)
    return idx


def sample_and_group(npoint: int, radius: float, nsample: int, xyz: torch.Tensor, points: torch.Tensor | None,
                     returnfps: bool = False, start_idx: torch.Tensor | None = None):
    """models/openshape/pointnet_util.py:113-146: FPS + ball query + gather + centre subtraction + concat."""
    B, N, _ = xyz.shape
    if start_idx is None:
        start_idx = torch.randint(0, N, (B,), dtype=torch.long)
    fps_idx, new_xyz = fps_sample(xyz, npoint, start_idx, want_idx=returnfps)
    idx, new_points = ball_group(xyz, new_xyz, radius, nsample, points, want_idx=returnfps)
    if returnfps:
        grouped_xyz = torch.gather(xyz.unsqueeze(1).expand(-1, npoint, -1, -1), 2,
                                   idx.unsqueeze(-1).expand(-1, -1, -1, 3))
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


class Group(nn.Module):
    """Uni3D / ULIP group divider. ``forward(xyz)`` (ULIP, dvae.py:159-181) returns (neighborhood, center);
    ``forward(xyz, color)`` (Uni3D, point_encoder.py:99-127) returns (neighborhood, center, features)."""

    def __init__(self, num_group: int, group_size: int, random_start: bool = False, skip_small_norm: bool = False,
                 device_rng: bool = False, pointnet2: bool | None = None):
        super().__init__()
        self.num_group = num_group
        self.group_size = group_size
        self.random_start = random_start      # True: ULIP (torch.randint start); False: Uni3D (pointnet2, start 0)
        self.skip_small_norm = skip_small_norm
        # Uni3D samples with the pointnet2_ops extension (models/point_encoder.py:12): its own arithmetic and tie order
        self.pointnet2 = (not random_start) if pointnet2 is None else pointnet2
        # False: draw the start on the global CPU generator exactly like the reference (misc.py:52);
        # True: draw it on the device generator (no host round-trip: required inside CUDA graphs)
        self.device_rng = device_rng
        self.next_start_idx: torch.Tensor | None = None  # one-shot override used by parity harnesses

    def forward(self, xyz: torch.Tensor, color: torch.Tensor | None = None):
        B, N, _ = xyz.shape
        start = None
        if self.next_start_idx is not None:
            start, self.next_start_idx = self.next_start_idx, None
        elif self.random_start:
            start = torch.randint(0, N, (B,), dtype=torch.long, device=xyz.device if self.device_rng else 'cpu')
        _, center = fps_sample(xyz, self.num_group, start, skip_small_norm=self.skip_small_norm, want_idx=False,
                               pointnet2=self.pointnet2 and start is None)
        _, neighborhood, features = knn_group(xyz, center, self.group_size, color)
        if color is None:
            return neighborhood, center
        return neighborhood, center, features
