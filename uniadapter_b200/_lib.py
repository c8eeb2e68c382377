"""ctypes binding of libua_b200.so (the C ABI declared in include/ua_b200.h).

The shared library is built in-tree by ``uniadapter_b200/csrc/Makefile`` (nvcc, sm_100a only). There is no CPU
fallback: every op of this package calls through this module and raises if the library is missing or a call
returns an error code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libua_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

UA_OK = 0
UA_ERR_UNSUPPORTED = -2
_lib = None

_P = C.c_void_p
_I = C.c_int
_F = C.c_float

# name -> (restype, argtypes); mirrors include/ua_b200.h one to one
SIGNATURES = {
    "ua_abi_version": (_I, []),
    "ua_last_error": (C.c_char_p, []),
    "ua_launch_count": (C.c_int64, []),
    "ua_reset_launch_count": (None, []),
    "ua_set_tuning": (_I, [C.c_char_p, _I]),
    "ua_fps_f32": (_I, [_P, _I, _I, _I, _P, _I, _P, _I, _P, _P, _P]),
    "ua_knn_group_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "ua_ball_group_f32": (_I, [_P, _P, _I, _P, _I, _I, _I, _F, _I, _P, _I, _P, _P]),
    "ua_gather_points_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "ua_head_f32": (_I, [_P, _I, _I, _P, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    "ua_head_prepare_f32": (_I, [_P, _I, _I, _F, _P, _P, _P, _P]),
    "ua_row_stats_f32": (_I, [_P, _I, _I, C.c_longlong, _P, _P, _P, _P]),
    "ua_modedota_step_f32": (_I, [_P, _I, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _I, _I, _P]),
    "ua_modedota_sample_step_f32": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _I, _I, _P]),
    "ua_debug_sample_trace": (_I, [_P]),
    "ua_modedota_sharded_step_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P]),
    "ua_fuse_logits_f32": (_I, [_P, _P, _I, _I, _I, _P, _I, _I, _F, _F, _F, _F, _F, _I, _P, _P, _P, _P]),
    "ua_stream_rng_f32": (_I, [_P, _P, _I, C.c_longlong, _P, _P, _I, _P, _P]),
    "ua_residual_scratch_floats": (C.c_longlong, [_I, _I, _I, _I]),
    "ua_residual_learn_f32": (_I, [_P, C.c_longlong, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, C.c_double,
                                    C.c_double, C.c_double, C.c_double, _I, _P, _P, _P, C.c_longlong, _P]),
    "ua_align_loss_grad_f32": (_I, [_P, C.c_longlong, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P,
                                     C.c_longlong, _P]),
    "ua_split_tf32_f32": (_I, [_P, _P, _P, C.c_longlong, _P]),
    "ua_pointwise_linear_split_f32": (_I, [_P, _P, _P, _I, C.c_longlong, _I, _I, _P, _P, _P]),
    "ua_gemm_tf32x3_f32": (_I, [_P, _P, C.c_longlong, _P, _P, C.c_longlong, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P,
                                 C.c_longlong, _P, _P, _P, _P]),
    "ua_layernorm_split_f32": (_I, [_P, _P, _P, _P, _F, C.c_longlong, _I, _P, _P, _P, _P]),
    "ua_attn_padded_tokens": (C.c_longlong, [_I]),
    "ua_attn_prepare_f32": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "ua_attention_f32": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "ua_dota_fit_f32": (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _I, _P]),
    "ua_dota_predict_f16": (_I, [_P, _I, _P, _P, _I, _I, _P, _P]),
    "ua_dota_regularize_f32": (_I, [_P, _I, _F, _P, _P]),
    "ua_dota_update_workspace_bytes": (C.c_longlong, [_I]),
    "ua_dota_update_f32": (_I, [_P, _I, _F, _P, _P, _P, _P]),
}


class ShardRank(C.Structure):
    """``ua_shard_rank`` of include/ua_b200.h (a host array; the library copies it into the kernel parameters)."""
    _fields_ = [("x_fit", _P), ("x_fit2", _P), ("text_local", _P), ("clip_local", _P), ("mu", _P), ("var", _P), ("pi", _P), ("c", _P),
                ("class_counts", _P), ("peer_recv", _P), ("peer_flag", _P), ("seq", _P), ("err", _P), ("done", _P),
                ("c_sum", _P), ("out_final", _P), ("out_argmax", _P), ("out_clip", _P), ("out_dota", _P), ("rank", _I),
                ("reserved", _I)]


class UaError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libua_b200.so in-tree (nvcc cross-compiles for sm_100a without a GPU)."""
    proc = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise UaError("building libua_b200.so failed:\n" + proc.stdout[-4000:] + proc.stderr[-4000:])
    if verbose:
        print(proc.stdout[-2000:])
    return LIB_PATH


class _TimedLib:
    """Proxy around the CDLL that brackets every kernel entry point with CUDA events (profiling aid for bench.py)."""

    def __init__(self, handle):
        self._h = handle
        self.records = {}
        self.flops = {}

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not name.startswith("ua_") or name in ("ua_last_error", "ua_launch_count", "ua_reset_launch_count",
                                                  "ua_abi_version", "ua_set_tuning"):
            return fn

        def timed(*args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            self.records.setdefault(name, []).append((s, e))
            if name == "ua_gemm_tf32x3_f32":       # M, N, K -> fp32-equivalent flops of this launch
                self.flops[name] = self.flops.get(name, 0) + 2 * int(args[6]) * int(args[7]) * int(args[8])
            return rc

        return timed

    def summary(self):
        """name -> (calls, total_us, mean_us); call after a synchronize."""
        out = {}
        for name, evs in self.records.items():
            us = [s.elapsed_time(e) * 1e3 for s, e in evs]
            out[name] = (len(us), sum(us), sum(us) / len(us))
        return out


_timed = None


def enable_kernel_timing(flag: bool = True):
    """Route calls through the event-timing proxy (eager mode only; not usable under CUDA-graph capture)."""
    global _timed
    _timed = _TimedLib(lib()) if flag else None
    return _timed


def lib():
    """The loaded library. Fails loudly when it has not been built: no other implementation exists."""
    global _lib
    if _timed is not None:
        return _timed
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UaError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` (or __graft_entry__.build()). "
                "uniadapter_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        for kv in filter(None, os.environ.get("UA_TUNING", "").split(",")):     # experiments: "key=value,key=value"
            key, _, val = kv.partition("=")
            check(handle.ua_set_tuning(key.strip().encode(), int(val)), f"UA_TUNING {kv}")
    return _lib


def ptr(t: torch.Tensor | None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise UaError("uniadapter_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise UaError("uniadapter_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def check(rc: int, what: str) -> None:
    if rc != UA_OK:
        raise UaError(f"{what} failed with code {rc}: {lib().ua_last_error().decode()}")


def launch_count() -> int:
    return int(lib().ua_launch_count())


def reset_launch_count() -> None:
    lib().ua_reset_launch_count()


def set_tuning(key: str, value: int) -> None:
    check(lib().ua_set_tuning(key.encode(), int(value)), "ua_set_tuning")
