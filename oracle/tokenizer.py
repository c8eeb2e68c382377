"""TEST INFRASTRUCTURE — ctypes wrapper of the C tokenizer oracle (tokenizer_oracle.c). See oracle/__init__.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_tokenizer.so")
_lib = None


def build() -> str:
    proc = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building the C oracle failed:\n" + proc.stdout + proc.stderr)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int64)
        _lib.oracle_fps.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, ip, C.c_int, ip]
        _lib.oracle_fps_pointnet2.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, ip]
        _lib.oracle_fps_pointnet2.restype = None
        _lib.oracle_knn.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, fp]
        _lib.oracle_ball.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, ip]
        _lib.oracle_sqdist.argtypes = [fp, C.c_int, fp, C.c_int, fp]
        for f in (_lib.oracle_fps, _lib.oracle_knn, _lib.oracle_ball, _lib.oracle_sqdist):
            f.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _ranges(n, parts):
    parts = max(1, min(parts, n))
    edges = np.linspace(0, n, parts + 1).astype(int)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(parts) if edges[i + 1] > edges[i]]


def fps(xyz: np.ndarray, npoint: int, start_idx=None, skip_small_norm: bool = False, threads: int = 1) -> np.ndarray:
    """xyz (B,N,3) f32 -> indices (B,npoint) int64. start_idx None -> 0 (pointnet2_ops convention)."""
    lib = _load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), dtype=np.int64)
    st = None if start_idx is None else np.ascontiguousarray(start_idx, dtype=np.int64)
    stp = _ip(st) if st is not None else None

    def run(r):
        lib.oracle_fps(_fp(xyz), r[0], r[1], N, npoint, stp, int(skip_small_norm), _ip(out))

    rs = _ranges(B, threads)
    if len(rs) == 1:
        run(rs[0])
    else:
        with ThreadPoolExecutor(len(rs)) as ex:
            list(ex.map(run, rs))
    return out


def fps_pointnet2(xyz: np.ndarray, npoint: int, threads: int = 1) -> np.ndarray:
    """The published pointnet2_ops furthest_point_sampling_kernel (Uni3D's FPS, models/point_encoder.py:7-14), thread by
    thread: xyz (B,N,3) f32 -> indices (B,npoint) int64, first sample 0."""
    lib = _load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), dtype=np.int64)

    def run(r):
        lib.oracle_fps_pointnet2(_fp(xyz), r[0], r[1], N, npoint, _ip(out))

    rs = _ranges(B, threads)
    if len(rs) == 1:
        run(rs[0])
    else:
        with ThreadPoolExecutor(len(rs)) as ex:
            list(ex.map(run, rs))
    return out


def knn(xyz: np.ndarray, centers: np.ndarray, k: int, threads: int = 1, return_dist: bool = False):
    """k nearest (expanded-form distance, ties -> lower index), nearest first. -> (B,G,k) int64 [, dist f32]."""
    lib = _load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    B, N, _ = xyz.shape
    G = centers.shape[1]
    out = np.empty((B, G, k), dtype=np.int64)
    dist = np.empty((B, G, k), dtype=np.float32) if return_dist else None

    def run(r):
        lib.oracle_knn(_fp(xyz), _fp(centers), B, N, G, r[0], r[1], k, _ip(out), _fp(dist) if return_dist else None)

    rs = _ranges(G, threads)
    if len(rs) == 1:
        run(rs[0])
    else:
        with ThreadPoolExecutor(len(rs)) as ex:
            list(ex.map(run, rs))
    return (out, dist) if return_dist else out


def radius2_f32(radius: float) -> float:
    """The fp32 value the reference's ``sqrdists > radius ** 2`` compares against (pointnet_util.py:105)."""
    return float(np.float32(float(radius) ** 2))


def ball(xyz: np.ndarray, centers: np.ndarray, radius: float, nsample: int, threads: int = 1) -> np.ndarray:
    lib = _load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    B, N, _ = xyz.shape
    S = centers.shape[1]
    out = np.empty((B, S, nsample), dtype=np.int64)

    def run(r):
        lib.oracle_ball(_fp(xyz), _fp(centers), B, N, S, r[0], r[1], radius2_f32(radius), nsample, _ip(out))

    rs = _ranges(S, threads)
    if len(rs) == 1:
        run(rs[0])
    else:
        with ThreadPoolExecutor(len(rs)) as ex:
            list(ex.map(run, rs))
    return out


def sqdist(centers: np.ndarray, xyz: np.ndarray) -> np.ndarray:
    """(G,3),(N,3) -> (G,N) expanded-form distances of one cloud."""
    lib = _load()
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    out = np.empty((centers.shape[0], xyz.shape[0]), dtype=np.float32)
    lib.oracle_sqdist(_fp(centers), centers.shape[0], _fp(xyz), xyz.shape[0], _fp(out))
    return out


def gather(xyz: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """index_points: (B,N,C), (B,...) -> (B,...,C)."""
    B = xyz.shape[0]
    return np.stack([xyz[b][idx[b]] for b in range(B)], axis=0)


def group_knn(xyz: np.ndarray, npoint: int, k: int, rgb=None, start_idx=None, skip_small_norm=False, threads: int = 1,
              sort_by_index: bool = False, pointnet2: bool = False):
    """Group.forward (point_encoder.py:99-127 / dvae.py:159-181): returns dict(fps_idx, center, idx, neigh[, feat]).
    Neighbours nearest-first, or in ascending point index (the CUDA kernel's emission order) if sort_by_index."""
    fidx = fps_pointnet2(xyz, npoint, threads) if pointnet2 else fps(xyz, npoint, start_idx, skip_small_norm, threads)
    center = gather(xyz, fidx)
    idx = knn(xyz, center, k, threads)
    if sort_by_index:
        idx = np.sort(idx, axis=-1)
    neigh = gather(xyz, idx) - center[:, :, None, :]
    out = dict(fps_idx=fidx, center=center, idx=idx, neigh=neigh.astype(np.float32))
    if rgb is not None:
        out["feat"] = np.concatenate([out["neigh"], gather(np.asarray(rgb, dtype=np.float32), idx)], axis=-1)
    return out


def sample_and_group(xyz: np.ndarray, npoint: int, radius: float, nsample: int, points=None, start_idx=None,
                     threads: int = 1):
    """pointnet_util.py:113-146: returns dict(fps_idx, new_xyz, idx, new_points)."""
    fidx = fps(xyz, npoint, start_idx, False, threads)
    new_xyz = gather(xyz, fidx)
    idx = ball(xyz, new_xyz, radius, nsample, threads)
    grouped = gather(xyz, idx) - new_xyz[:, :, None, :]
    new_points = grouped.astype(np.float32)
    if points is not None:
        new_points = np.concatenate([new_points, gather(np.asarray(points, dtype=np.float32), idx)], axis=-1)
    return dict(fps_idx=fidx, new_xyz=new_xyz, idx=idx, new_points=new_points)
