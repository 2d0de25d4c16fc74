"""ctypes binding of libslime_b200.so (C ABI: include/slime_b200.h).

The library is built in-tree by slime_mold_b200/build.py.  Loading never falls back
to anything else: if the shared object is missing or the machine has no sm_100
device, calls raise `SlimeError`.
"""
from __future__ import annotations

import ctypes as C
import os

from .settings import SimSizeUniform

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SM_LIB_PATH") or os.path.join(_HERE, "libslime_b200.so")   # SM_LIB_PATH: A/B builds only

SM_FLAG_GAUSSIAN_BLUR = 1 << 0
SM_FLAG_NO_SORT = 1 << 1
SM_COMM_ID_BYTES = 128

STATUS_NAMES = {0: "SM_OK", -1: "SM_ERR_BAD_ARG", -2: "SM_ERR_CUDA", -3: "SM_ERR_NCCL", -4: "SM_ERR_OOM",
                -5: "SM_ERR_NO_DEVICE", -6: "SM_ERR_STATE"}


class SlimeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


class SmConfig(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("agent_count", C.c_uint64),
        ("device", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32),
        ("flags", C.c_uint32), ("sort_interval", C.c_uint32), ("reserved", C.c_uint32),
    ]


class SmTiming(C.Structure):
    _fields_ = [
        ("agents_ms", C.c_double), ("trail_ms", C.c_double), ("sort_ms", C.c_double), ("exchange_ms", C.c_double),
        ("agent_launches", C.c_uint64), ("trail_launches", C.c_uint64), ("sort_launches", C.c_uint64),
        ("exchange_launches", C.c_uint64), ("steps", C.c_uint64), ("kernel_launches", C.c_uint64),
    ]


class SmTrailStats(C.Structure):
    _fields_ = [("sum", C.c_double), ("sum_sq", C.c_double), ("max", C.c_float), ("_pad", C.c_uint32),
                ("nonzero", C.c_uint64)]


# every symbol include/slime_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
_E = C.c_void_p
SIGNATURES = {
    "sm_last_error": (C.c_char_p, []),
    "sm_version": (None, [_P(C.c_int), _P(C.c_int)]),
    "sm_device_count": (C.c_int, []),
    "sm_create": (C.c_int, [_P(_E), _P(SmConfig)]),
    "sm_destroy": (C.c_int, [_E]),
    "sm_comm_unique_id": (C.c_int, [_P(C.c_uint8)]),
    "sm_comm_init": (C.c_int, [_E, _P(C.c_uint8)]),
    "sm_set_params": (C.c_int, [_E, _P(SimSizeUniform)]),
    "sm_get_params": (C.c_int, [_E, _P(SimSizeUniform)]),
    "sm_upload_agents": (C.c_int, [_E, _P(C.c_float), C.c_uint64, C.c_uint64]),
    "sm_download_agents": (C.c_int, [_E, _P(C.c_float), C.c_uint64, C.c_uint64, _P(C.c_uint64)]),
    "sm_init_agents": (C.c_int, [_E, C.c_uint64]),
    "sm_set_agent_count": (C.c_int, [_E, C.c_uint64, C.c_uint64]),
    "sm_reassign_speeds": (C.c_int, [_E, C.c_uint64]),
    "sm_agent_count": (C.c_uint64, [_E]),
    "sm_local_agent_count": (C.c_uint64, [_E]),
    "sm_clear_trail": (C.c_int, [_E]),
    "sm_upload_trail": (C.c_int, [_E, _P(C.c_float), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t]),
    "sm_download_trail": (C.c_int, [_E, _P(C.c_float), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t]),
    "sm_trail_statistics": (C.c_int, [_E, _P(SmTrailStats)]),
    "sm_resize": (C.c_int, [_E, C.c_uint32, C.c_uint32]),
    "sm_set_lut": (C.c_int, [_E, _P(C.c_uint8)]),
    "sm_render_rgba8": (C.c_int, [_E, C.c_uint32, C.c_uint32, _P(C.c_uint8)]),
    "sm_save_snapshot": (C.c_int, [_E, C.c_char_p]),
    "sm_load_snapshot": (C.c_int, [_E, C.c_char_p]),
    "sm_step": (C.c_int, [_E, C.c_uint32]),
    "sm_diffuse_only": (C.c_int, [_E, C.c_uint32]),
    "sm_sync": (C.c_int, [_E]),
    "sm_get_timing": (C.c_int, [_E, _P(SmTiming)]),
    "sm_reset_timing": (C.c_int, [_E]),
    "sm_set_timing_enabled": (C.c_int, [_E, C.c_int]),
    "sm_stream": (C.c_void_p, [_E]),
    "sm_test_math": (C.c_int, [C.c_int, C.c_int, _P(C.c_float), _P(C.c_float), _P(C.c_int32), _P(C.c_float),
                               _P(C.c_float), C.c_uint64]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and attach signatures.  No fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SlimeError(-5, f"{LIB_PATH} is missing: run `python -m slime_mold_b200.build` "
                             "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sm_last_error()
        raise SlimeError(rc, msg.decode() if msg else "")
