"""ctypes binding of ``include/mixgrpo_b200.h`` — the only way Python reaches the CUDA kernels.

There is NO fallback: if the shared library is missing and cannot be built, or a call returns a
non-zero code, a ``RuntimeError`` is raised.  Nothing in here imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import _build

F32, BF16 = 0, 1
SRC_NOISE, SRC_GIVEN, SRC_DETERMINISTIC, SRC_PHILOX = 0, 1, 2, 3
FLAG_ROUND_LIKE_TORCH = 1
FLAG_PDL_EARLY_LOADS = 2
FLAG_PDL_EARLY_V = 4
FLAG_DEFER_LOGP = 8
POLICY_MAX_ITEMS = 8
ADV_GROUP_LOCAL, ADV_GROUP_SPLIT, ADV_GLOBAL = 0, 1, 2
EUNSUPPORTED = -4
PEER_MAX_WORLD, PEER_HANDLE_BYTES = 16, 64
ABI_VERSION = 6


class StepCoefs(C.Structure):
    """mirror of ``mixgrpo_step_coefs`` (include/mixgrpo_b200.h)."""
    _fields_ = [("two_var", C.c_float), ("log_scale", C.c_float), ("log_norm", C.c_float), ("c", C.c_float * 16)]


class PhiloxArgs(C.Structure):
    """mirror of ``mixgrpo_philox_args`` (include/mixgrpo_b200.h)."""
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint64), ("device_state", C.c_void_p)]


class StepExt(C.Structure):
    """mirror of ``mixgrpo_step_ext`` (include/mixgrpo_b200.h): the decode-ready second output of a step launch."""
    _fields_ = [("decode_out", C.c_void_p), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("divisor", C.c_float),
                ("shift", C.c_float), ("from_x0", C.c_int), ("reciprocal", C.c_int), ("x_is_bf16", C.c_int), ("x_f32_out", C.c_void_p),
                ("x_f32_out_bs", C.c_int64)]


class LossArgs(C.Structure):
    """mirror of ``mixgrpo_loss_args`` (include/mixgrpo_b200.h)."""
    _fields_ = [("old_logp", C.c_void_p), ("advantages", C.c_void_p), ("stats_rows", C.c_void_p),
                ("clip_range", C.c_double), ("adv_clip_max", C.c_double), ("kl_coeff", C.c_double), ("denom", C.c_double),
                ("accumulate", C.c_int)]


class PolicyItem(C.Structure):
    """mirror of ``mixgrpo_policy_item`` (include/mixgrpo_b200.h)."""
    _fields_ = [("v", C.c_void_p), ("x", C.c_void_p), ("x_next", C.c_void_p), ("x_bs", C.c_int64), ("in_bs", C.c_int64),
                ("logp", C.c_void_p), ("old_logp", C.c_void_p), ("stats_rows", C.c_void_p), ("grad_v", C.c_void_p),
                ("coefs", StepCoefs)]


_P, _I64, _I, _U, _F, _D = C.c_void_p, C.c_int64, C.c_int, C.c_uint, C.c_float, C.c_double
_CP = C.POINTER(StepCoefs)
_LP = C.POINTER(LossArgs)
_XP = C.POINTER(StepExt)
_IP = C.POINTER(PolicyItem)

# name -> (restype, argtypes); every symbol include/mixgrpo_b200.h declares
SIGNATURES = {
    "mixgrpo_step_workspace_bytes": (_I64, [_I64, _I64]),
    "mixgrpo_deferred_workspace_bytes": (_I64, [_I64, _I64]),
    "mixgrpo_abi_version": (_I, []),
    "mixgrpo_build_info": (C.c_char_p, []),
    "mixgrpo_set_tuning": (_I, [_I, _I]),
    "mixgrpo_error_string": (C.c_char_p, [_I]),
    "mixgrpo_flow_step": (_I, [_P, _I, _P, _I64, _P, _P, _I64, _P, _I64, _P, _P, _P, _P, _I64, _I64, _I64, _CP, _I, _U, _P, _XP]),
    "mixgrpo_dance_step": (_I, [_P, _I, _P, _I64, _P, _P, _I64, _P, _I64, _P, _P, _P, _P, _I64, _I64, _I64, _CP, _I, _I, _U, _P, _XP]),
    "mixgrpo_dpm_step": (_I, [_P, _I, _P, _I64, _P, _P, _P, _I, _P, _I64, _P, _P, _P, _P, _I64, _I64, _I64, _CP, _I, _U, _P, _XP]),
    "mixgrpo_logp_finalize": (_I, [_P, _I64, _I64, _I64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), _P, _I64, _P]),
    "mixgrpo_philox_advance": (_I, [_P, C.c_uint64, _P]),
    "mixgrpo_logprob_bwd": (_I, [_I, _P, _I, _P, _I64, _P, _I64, _P, _P, _I64, _I64, _CP, _U, _P]),
    "mixgrpo_policy_fwd": (_I, [_I, _P, _I, _P, _I64, _P, _I64, _P, _P, _I64, _I64, _I64, _CP, _LP, _U, _P]),
    "mixgrpo_policy_bwd": (_I, [_I, _P, _I, _P, _I64, _P, _I64, _P, _LP, _P, _I64, _I64, _CP, _U, _P]),
    "mixgrpo_policy_fwd_multi": (_I, [_I, _I, _IP, _I, _P, _D, _D, _D, _D, _I, _P, _I64, _I64, _I64, _U, _P]),
    "mixgrpo_policy_bwd_multi": (_I, [_I, _I, _IP, _I, _P, _D, _D, _D, _D, _I64, _I64, _U, _P]),
    "mixgrpo_policy_step": (_I, [_I, _P, _I, _P, _I64, _P, _I64, _P, _P, _P, _I64, _I64, _I64, _CP, _LP, _U, _P]),
    "mixgrpo_cast_rows": (_I, [_P, _I, _P, _I64, _I64, _I64, _P]),
    "mixgrpo_group_advantages": (_I, [_P, _P, _I, _I64, _I, _I, _I, _P, _I64, _P, _P]),
    "mixgrpo_grpo_loss": (_I, [_P, _P, _P, _I64, _D, _D, _D, _D, _P, _P, _P, _P]),
    "mixgrpo_pack_latents": (_I, [_P, _P, _I, _I64, _I, _I, _I, _P]),
    "mixgrpo_unpack_latents": (_I, [_P, _P, _I, _I64, _I, _I, _I, _F, _F, _P]),
    "mixgrpo_peer_region_bytes": (_I64, [_I, _I64]),
    "mixgrpo_peer_region_alloc": (_I, [_I, _I64, C.POINTER(_P), _P]),
    "mixgrpo_peer_region_open": (_I, [_P, C.POINTER(_P)]),
    "mixgrpo_peer_region_close": (_I, [_P]),
    "mixgrpo_peer_region_free": (_I, [_P]),
    "mixgrpo_peer_region_status": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "mixgrpo_peer_gather_advantages": (_I, [C.POINTER(_P), _I, _I, _I64, _P, _P, _I, _I64, _I, _I, _I, _P, _P, _P]),
    "mixgrpo_peer_allreduce": (_I, [C.POINTER(_P), _I, _I, _I64, _P, _I, _I, _P]),
}

_lib: Optional[C.CDLL] = None


def library_path() -> str:
    return str(_build.LIB)


def lib() -> C.CDLL:
    """Load (building first if the in-tree .so is missing or stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _build.LIB.exists() or (not _build.is_fresh() and os.environ.get("MIXGRPO_NO_REBUILD") != "1"):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            if not _build.LIB.exists():
                raise RuntimeError(
                    f"mixgrpo_b200: CUDA library {_build.LIB} is missing and could not be built ({e}). "
                    "There is no CPU fallback; run `python -c 'import __graft_entry__ as g; g.build()'`.") from e
    h = C.CDLL(str(_build.LIB))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(h, name)          # AttributeError here = the .so does not export the header's symbol
        fn.restype, fn.argtypes = res, args
    if h.mixgrpo_abi_version() != ABI_VERSION:
        raise RuntimeError(f"mixgrpo_b200: ABI mismatch (lib {h.mixgrpo_abi_version()} != binding {ABI_VERSION})")
    _lib = h
    return h


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mixgrpo_error_string(rc).decode()
        raise RuntimeError(f"mixgrpo_b200: {what} failed with code {rc}: {msg}")
