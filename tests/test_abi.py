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


def test_residual_learning_plans_for_any_class_count(libpath):
    """Host logic only (no launch): the scratch planner of ua_residual_learn_f32 accepts the class counts of every BASELINE
    data set -- ModelNet40 / ScanObjectNN / ShapeNet-C, OmniObject3D's 216 (likelihood matrix out of shared memory,
    column-chunked backward) and Objaverse-LVIS's 1156 (P in the scratch too: its K*K floats show up in the plan) -- and
    still refuses the shapes it cannot tile."""
    from uniadapter_b200 import _lib
    f = _lib.lib().ua_residual_scratch_floats
    small = f(1, 40, 8, 512)
    assert small > 0
    assert f(15, 40, 8, 512) > small
    assert f(1, 216, 8, 512) > 0
    lvis = f(1, 1156, 8, 1024)
    assert lvis >= 1156 * 1156 * (8 + 2) + 1156 * 8 * 1024      # Wt + LM + P (global) + 1/v
    assert f(1, 40, 8, 520) == -1          # D % 128 != 0
    assert f(1, 40, 6, 512) == -1          # M not a multiple of 4


def test_tuning_knobs_are_validated(libpath):
    from uniadapter_b200 import _lib
    for key in ("knn_hist", "dota_ka", "dota_pdl", "sample_per", "sample_v", "fps_cluster"):
        _lib.set_tuning(key, 0)
    _lib.set_tuning("dota_pdl", 1)
    with pytest.raises(_lib.UaError):
        _lib.set_tuning("no_such_knob", 1)
