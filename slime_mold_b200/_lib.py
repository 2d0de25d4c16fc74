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
SM_FLAG_SEM_INPLACE = 1 << 2
SM_COMM_ID_BYTES = 128

STATUS_NAMES = {0: "SM_OK", -1: "SM_ERR_BAD_ARG", -2: "SM_ERR_CUDA", -3: "SM_ERR_NCCL", -4: "SM_ERR_OOM",
                -5: "SM_ERR_NO_DEVICE", -6: "SM_ERR_STATE"}


class SlimeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


class SmTuning(C.Structure):
    """`sm_tuning` of include/slime_b200.h: measurement switches, all zero = the engine's defaults."""

    _fields_ = [
        ("sampler", C.c_uint32), ("tile_shift_x", C.c_uint32), ("tile_shift_y", C.c_uint32),
        ("trail_rows_per_chunk", C.c_uint32), ("deposit_counts_only", C.c_uint32), ("generic_trail_kernel", C.c_uint32),
        ("surface_row_writes", C.c_uint32), ("no_step_graph", C.c_uint32),
        ("gauss_kernel", C.c_uint32), ("gauss_rows_max_radius", C.c_uint32), ("gauss_rows_packing", C.c_uint32),
        ("gauss_chunk_rows", C.c_uint32),
        ("exchange", C.c_uint32), ("serial_exchange", C.c_uint32), ("migrate_capacity", C.c_uint32), ("barrier_fence", C.c_uint32),
        ("debug_single_rank_strip", C.c_uint32), ("debug_side_timing", C.c_uint32),
        ("no_boundary_first", C.c_uint32),
        ("deposit_flag_layout", C.c_uint32),
        ("reserved", C.c_uint32 * 4),
    ]


class SmConfig(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("agent_count", C.c_uint64),
        ("device", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32),
        ("flags", C.c_uint32), ("sort_interval", C.c_uint32), ("ghost_rows", C.c_uint32),
        ("tuning", SmTuning),
    ]


_GAUSS_KERNELS = {"": 0, "auto": 0, "rows": 1, "stream": 2, "tile": 3, "two_pass": 4}


def tuning_from_env(env=None) -> SmTuning:
    """The test / bench harness convention: SM_* environment variables select A/B variants.  They are read HERE, by the
    harness layer, and handed to the library as an explicit `sm_tuning`; libslime_b200.so itself reads no environment
    variable for them.  Unset = 0 = default."""
    env = os.environ if env is None else env

    def num(name, dflt=0):
        v = env.get(name, "")
        return int(v) if v.strip() else dflt

    t = SmTuning()
    t.sampler = 1 if env.get("SM_SAMPLER", "") == "ldg" else 0
    t.tile_shift_x, t.tile_shift_y = num("SM_TILE_SHIFT_X"), num("SM_TILE_SHIFT_Y")
    t.trail_rows_per_chunk = num("SM_TRAIL_ROWS_PER_CHUNK")
    t.deposit_counts_only = num("SM_NO_DEPOSIT_FLAGS")
    t.generic_trail_kernel = num("SM_FORCE_GENERIC_TRAIL")
    t.surface_row_writes = 0 if num("SM_SURF_PAIRS", 1) else 1
    t.no_step_graph = 0 if num("SM_STEP_GRAPH", 1) else 1
    gk = env.get("SM_GAUSS_KERNEL", "")
    if gk not in _GAUSS_KERNELS:
        raise ValueError(f"SM_GAUSS_KERNEL={gk!r}: one of {sorted(_GAUSS_KERNELS)}")
    t.gauss_kernel = 4 if num("SM_GAUSS_TWO_PASS") else _GAUSS_KERNELS[gk]
    t.gauss_rows_max_radius = num("SM_GAUSS_ROWS_MAX_R")
    rp = num("SM_GAUSS_ROWS_PACKED", -1)               # -1 auto, 0 scalar taps, >= 1 FFMA2 taps
    t.gauss_rows_packing = 0 if rp < 0 else (1 if rp == 0 else 2)
    t.gauss_chunk_rows = num("SM_GAUSS_CHUNK")
    t.exchange = 1 if env.get("SM_EXCHANGE", "") == "nccl" else 0
    t.serial_exchange = 0 if num("SM_OVERLAP", 1) else 1
    t.migrate_capacity = num("SM_MIGRATE_CAP")
    bf = num("SM_BARRIER_FENCE", 1)                     # historical numbering: 0 sc fences, 1 acq_rel (default), 2 device scope
    t.barrier_fence = {0: 1, 1: 0, 2: 2}.get(bf, 0)
    t.debug_single_rank_strip = num("SM_FAKE_MULTI")
    t.debug_side_timing = num("SM_SIDE_TIMING")
    t.no_boundary_first = 0 if num("SM_BOUNDARY_FIRST", 1) else 1
    t.deposit_flag_layout = {"": 0, "auto": 0, "linear": 1, "tiled": 2}[env.get("SM_FLAG_LAYOUT", "")]
    return t


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
