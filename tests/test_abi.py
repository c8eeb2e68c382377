"""CPU: the C-ABI library builds, loads and exports every symbol include/ua_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ua_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ua_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def libpath():
    from uniadapter_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.LIB_PATH


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ["ua_fps_f32", "ua_knn_group_f32", "ua_ball_group_f32", "ua_head_f32", "ua_modedota_step_f32",
                 "ua_fuse_logits_f32", "ua_dota_fit_f32", "ua_dota_predict_f16"]:
        assert must in syms


def test_library_exports_every_declared_symbol(libpath):
    handle = ctypes.CDLL(libpath)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"declared in ua_b200.h but not exported: {missing}"


def test_binding_table_matches_header(libpath):
    from uniadapter_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().ua_abi_version() == 1


def test_no_cpu_fallback():
    """Ops must refuse CPU tensors instead of silently computing somewhere else."""
    import torch
    import uniadapter_b200 as ua
    with pytest.raises(ua._lib.UaError):
        ua.fps_sample(torch.zeros(1, 16, 3), 4)
    with pytest.raises(ua._lib.UaError):
        ua.zero_shot_head(torch.zeros(1, 8), torch.zeros(3, 8))
