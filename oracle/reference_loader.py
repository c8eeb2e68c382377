"""TEST INFRASTRUCTURE — imports the reference's own Python modules from /root/reference (build container only).

Used by oracle/make_golden.py to mint the golden vectors under tests/golden/ and to pin the oracle. Nothing at
run time on the GPU box may use this (the reference is absent there). Missing third-party packages are stubbed
the way SURVEY §8c describes; no reference source is copied.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # oracle/fetch_ref.py (git-ignored)
REF_ROOT = os.environ.get("UA_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference") else _LOCAL_COPY)


def available() -> bool:
    return os.path.isdir(REF_ROOT) and os.path.exists(os.path.join(REF_ROOT, "dota_mixture.py"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _load(path: str, modname: str, package: str | None = None):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, path))
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def install_stubs():
    import torch.nn as nn
    _stub("clip")
    _stub("open_clip")
    _stub("plotly")
    _stub("plotly.graph_objects")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("mpl_toolkits")
    _stub("mpl_toolkits.mplot3d", Axes3D=object)
    _stub("timm")
    _stub("timm.models")

    class DropPath(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()

    _stub("timm.models.layers", DropPath=DropPath)
    _stub("pointnet2_ops", pointnet2_utils=types.SimpleNamespace())


def dota_mixture():
    return _load("dota_mixture.py", "ref_dota_mixture")


def dota():
    install_stubs()
    return _load("dota.py", "ref_dota")


def uni_adapter():
    """Uni_Adapter.py (get_logits_wrapper, softmax_entropy, compute_text_alignment_loss)."""
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return _load("Uni_Adapter.py", "ref_uni_adapter")


def ulip_pointbert():
    """(misc, dvae, point_encoder) of models/ulip/pointbert loaded under a fake package."""
    install_stubs()
    pkg = "ref_ulip_pointbert"
    if pkg not in sys.modules:
        p = types.ModuleType(pkg)
        p.__path__ = [os.path.join(REF_ROOT, "models/ulip/pointbert")]
        sys.modules[pkg] = p
    misc = _load("models/ulip/pointbert/misc.py", pkg + ".misc", pkg)
    dvae = _load("models/ulip/pointbert/dvae.py", pkg + ".dvae", pkg)
    penc = _load("models/ulip/pointbert/point_encoder.py", pkg + ".point_encoder", pkg)
    return misc, dvae, penc


def openshape_pointnet_util():
    return _load("models/openshape/pointnet_util.py", "ref_openshape_pointnet_util")


def _install_redstone_stub():
    """torch_redstone is absent; ppta.py uses two of its helpers: ``Lambda`` (a module around a function) and
    ``supercat`` (concatenate after broadcasting the other dimensions)."""
    import torch
    import torch.nn as nn

    class Lambda(nn.Module):
        def __init__(self, fn):
            super().__init__()
            self.fn = fn

        def forward(self, *a, **k):
            return self.fn(*a, **k)

    def supercat(tensors, dim=0):
        nd = max(t.dim() for t in tensors)
        ts = [t.reshape((1,) * (nd - t.dim()) + tuple(t.shape)) for t in tensors]
        d = dim % nd
        shape = [max(t.shape[i] for t in ts) for i in range(nd)]
        out = []
        for t in ts:
            tgt = list(shape)
            tgt[d] = t.shape[d]
            out.append(t.expand(*tgt))
        return torch.cat(out, dim=d)

    _stub("torch_redstone", Lambda=Lambda, supercat=supercat)


def openshape_ppta():
    """models/openshape/{pointnet_util,ppta}.py under a fake package (the real __init__ pulls more dependencies)."""
    install_stubs()
    _install_redstone_stub()
    pkg = "ref_openshape"
    if pkg not in sys.modules:
        p = types.ModuleType(pkg)
        p.__path__ = [os.path.join(REF_ROOT, "models/openshape")]
        sys.modules[pkg] = p
    pu = _load("models/openshape/pointnet_util.py", pkg + ".pointnet_util", pkg)
    ppta = _load("models/openshape/ppta.py", pkg + ".ppta", pkg)
    return pu, ppta


def uni3d_point_encoder():
    """models/point_encoder.py with pointnet2_ops stubbed (its fps() is NOT runnable: un-vendored CUDA)."""
    install_stubs()
    return _load("models/point_encoder.py", "ref_uni3d_point_encoder")
