"""TEST INFRASTRUCTURE — the parity cases (shapes + seeds) and their deterministic inputs.

Shared by oracle/make_golden.py (which runs the reference on them) and by the tests (which run the oracle and the
CUDA path on them). Shapes follow BASELINE.json's configs plus ragged / tiny / tie-heavy cases.
"""
from __future__ import annotations

import numpy as np

from . import synth

CFG = {'epsilon': 1e-4, 'sigma': 1e-4, 'eta': 0.1, 'rho': 0.02}   # utils/params.py:103-106 of the reference

# name -> (B, N, G, k, seed, start mode)
TOK_KNN = {
    "tok_ulip_b2_n1024_g512_k32": (2, 1024, 512, 32, 11, "random"),      # cfg 1/2 (ULIP-2 PointBERT)
    "tok_uni3d_b1_n10000_g512_k64": (1, 10000, 512, 64, 12, "zero"),     # cfg 4 (Uni3D-L, 10k points)
    "tok_uni3d_b3_n1024_g512_k64": (3, 1024, 512, 64, 13, "zero"),       # cfg 5 shape
    "tok_ragged_b3_n257_g40_k9": (3, 257, 40, 9, 14, "random"),          # ragged sizes
    "tok_tiny_b2_n33_g33_k33": (2, 33, 33, 33, 15, "random"),            # G = k = N
    "tok_dups_b1_n300_g64_k16": (1, 300, 64, 16, 16, "random"),          # every point three times: ties everywhere
}

# name -> (B, N, S, radius, nsample, seed)
TOK_BALL = {
    "tok_oshape_b1_n10000_s384_r02_ns64": (1, 10000, 384, 0.2, 64, 21),  # cfg 3 (OpenShape PPAT scaling 4)
    "tok_oshape_b2_n700_s50_r015_ns16": (2, 700, 50, 0.15, 16, 22),
    "tok_oshape_sparse_b1_n200_s20_r005_ns32": (1, 200, 20, 0.05, 32, 23),   # balls with < nsample hits -> padding
}

# name -> (B, D, K, seed)
HEAD = {
    "head_b1_d512_k40": (1, 512, 40, 31),
    "head_b1_d1024_k1156": (1, 1024, 1156, 32),
    "head_b64_d1024_k55": (64, 1024, 55, 33),
    "head_b5_d1280_k15": (5, 1280, 15, 34),
}

# name -> (K, M, D, B, T, seed, store full state)
MODEDOTA = {
    "modedota_k40_m8_d512_b1": (40, 8, 512, 1, 12, 41, False),       # cfg 2
    "modedota_k10_m4_d128_b1": (10, 4, 128, 1, 30, 42, True),
    "modedota_k15_m8_d1280_b1": (15, 8, 1280, 1, 6, 43, False),      # cfg 3
    "modedota_k55_m8_d1024_b64": (55, 8, 1024, 64, 3, 44, False),    # cfg 5
    "modedota_k7_m3_d50_b5": (7, 3, 50, 5, 8, 45, True),             # odd sizes (no 16-byte tiles)
}

# name -> (K, D, B, T, seed)
DOTA = {
    "dota_k40_d128_b1": (40, 128, 1, 10, 51),
    "dota_k12_d64_b4": (12, 64, 4, 6, 52),
}

# name -> (K, M, D, seed)
ALIGN = {
    "align_k10_m4_d64": (10, 4, 64, 61),
    "align_k40_m8_d512": (40, 8, 512, 62),
    "align_k15_m4_d128": (15, 4, 128, 63),
}


# name -> (T steps, N points, K classes, M modes, encoder depth, res_learning, seed)
E2E = {
    "e2e_ulip_d2_modedota_res": (8, 1024, 40, 8, 2, True, 71),
    "e2e_ulip_d2_modedota": (8, 1024, 40, 8, 2, False, 72),
    "e2e_ulip_d2_dota": (6, 1024, 40, 0, 2, False, 73),
}
# OpenShape (cfg 3): name -> (T steps, N points, S patches, K classes, M modes, depth, seed); D = 1280, rgb ~ U[0,1)
E2E_OSHAPE = {
    "e2e_oshape_d2_modedota": (5, 10000, 384, 15, 8, 2, 81),
}
# Uni3D front end (cfg 4/5 geometry): name -> (B, N, G, k, encoder dim, trans dim, seed)
UNI3D_FRONT = {
    "uni3d_front_b1_n10000": (1, 10000, 512, 64, 512, 1024, 91),
    "uni3d_front_b2_n1024": (2, 1024, 512, 64, 512, 1024, 92),
}
E2E_MODEL_SEED = 0
E2E_LOOP_SEED = 123


def e2e_inputs(name):
    T, N, K, M, depth, res, seed = E2E[name]
    return dict(pc=synth.cloud(T, N, seed), text=synth.unit_rows(K, 512, seed + 1), T=T, N=N, K=K, M=M, depth=depth,
                res_learning=res)


def e2e_oshape_inputs(name):
    T, N, S, K, M, depth, seed = E2E_OSHAPE[name]
    return dict(pc=synth.cloud(T, N, seed), rgb=synth.uniform((T, N, 3), seed + 2), text=synth.unit_rows(K, 1280, seed + 1),
                T=T, N=N, S=S, K=K, M=M, depth=depth)


def uni3d_front_inputs(name):
    B, N, G, k, enc, trans, seed = UNI3D_FRONT[name]
    return dict(xyz=synth.cloud(B, N, seed), rgb=synth.uniform((B, N, 3), seed + 1), B=B, N=N, G=G, k=k, enc=enc, trans=trans)


def tok_knn_inputs(name):
    B, N, G, k, seed, mode = TOK_KNN[name]
    xyz = synth.cloud(B, N, seed)
    if "dups" in name:
        xyz[:, 100:200] = xyz[:, 0:100]
        xyz[:, 200:300] = xyz[:, 0:100]
    rgb = synth.uniform((B, N, 3), seed + 1000)
    start = synth.integers(0, N, (B,), seed + 2000) if mode == "random" else np.zeros(B, dtype=np.int64)
    return dict(xyz=xyz, rgb=rgb, start=start, B=B, N=N, G=G, k=k)


def tok_ball_inputs(name):
    B, N, S, radius, nsample, seed = TOK_BALL[name]
    xyz = synth.cloud(B, N, seed)
    rgb = synth.uniform((B, N, 3), seed + 1000)
    points = np.concatenate([xyz, rgb], axis=-1)
    start = synth.integers(0, N, (B,), seed + 2000)
    return dict(xyz=xyz, points=points, start=start, B=B, N=N, S=S, radius=radius, nsample=nsample)


def head_inputs(name):
    B, D, K, seed = HEAD[name]
    x = (synth._rng(seed).standard_normal((B, D), dtype=np.float32) * np.float32(3.0) + np.float32(0.1)).astype(np.float32)
    return dict(x=x, text=synth.unit_rows(K, D, seed + 1), B=B, D=D, K=K)


def modedota_inputs(name):
    K, M, D, B, T, seed, full = MODEDOTA[name]
    text = synth.unit_rows(K, D, seed)
    x, xa, _ = synth.features(T, B, D, text, seed + 1)
    return dict(text=text, x=x, x_aug=xa, K=K, M=M, D=D, B=B, T=T, full=full)


def dota_inputs(name):
    K, D, B, T, seed = DOTA[name]
    text = synth.unit_rows(K, D, seed)
    x, _, _ = synth.features(T, B, D, text, seed + 1)
    return dict(text=text, x=x, K=K, D=D, B=B, T=T)


def align_inputs(name):
    K, M, D, seed = ALIGN[name]
    text = synth.unit_rows(K, D, seed)
    x, xa, _ = synth.features(5, 1, D, text, seed + 1)
    res = (np.float32(0.01) * synth._rng(seed + 2).standard_normal((K, D), dtype=np.float32) / np.float32(np.sqrt(D))).astype(np.float32)
    return dict(text=text, x=x, x_aug=xa, residual=res, K=K, M=M, D=D)


def input_crc(inp):
    return synth.crc(*[v for v in inp.values() if isinstance(v, np.ndarray)])
