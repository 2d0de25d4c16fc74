"""CudaBackend -- host-side mirror of the reference's compute plumbing.

It stands where /root/reference/src/pipeline_manager.rs (compute / decay / diffuse
pipelines) and /root/reference/src/bind_group_manager.rs (agents + trail + uniform bind
group) stand in the reference, and exposes the operations src/main.rs performs on
them: write the uniform (main.rs:83-99), fill / read back / rewrite the agent buffer
(:101-145, 263-282, 682-791, 954-997), clear or replace the trail (:909-913,
999-1015) and run one frame of compute passes (:1163-1235).  All device work
happens in libslime_b200.so through the C ABI of include/slime_b200.h.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _lib
from ._lib import SlimeError, SmConfig, SmTiming, SmTrailStats, SmTuning, check, tuning_from_env
from .settings import Settings, SimSizeUniform


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class CudaBackend:
    def __init__(self, width: int, height: int, settings: Optional[Settings] = None, *, agent_count: Optional[int] = None,
                 device: int = 0, rank: int = 0, world_size: int = 1, flags: int = 0, sort_interval: int = 0,
                 ghost_rows: int = 0, tuning: Optional[SmTuning] = None):
        """`tuning`: an explicit `sm_tuning` (measurement switches); None = what the SM_* environment variables of the test /
        bench harness say (`_lib.tuning_from_env`), all zero when none is set."""
        self._lib = _lib.load()
        self.settings = settings.clone() if settings is not None else Settings.default()
        if agent_count is not None:
            self.settings.agent_count = int(agent_count)
        self.width, self.height = int(width), int(height)
        self.rank, self.world_size = rank, world_size
        cfg = SmConfig(width=self.width, height=self.height, agent_count=self.settings.agent_count, device=device,
                       rank=rank, world_size=world_size, flags=flags, sort_interval=sort_interval, ghost_rows=ghost_rows,
                       tuning=tuning if tuning is not None else tuning_from_env())
        self._h = C.c_void_p()
        check(self._lib.sm_create(C.byref(self._h), C.byref(cfg)))
        self.update_settings(self.settings)

    # -- construction / teardown ------------------------------------------------
    @classmethod
    def new(cls, width: int, height: int, settings: Optional[Settings] = None, **kw) -> "CudaBackend":
        """PipelineManager::new + BindGroupManager::new + buffer creation (main.rs:263-293, 326-368)."""
        return cls(width, height, settings, **kw)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.sm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- multi-GPU ---------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * _lib.SM_COMM_ID_BYTES)()
        check(self._lib.sm_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes) -> None:
        assert len(unique_id) == _lib.SM_COMM_ID_BYTES
        buf = (C.c_uint8 * _lib.SM_COMM_ID_BYTES).from_buffer_copy(unique_id)
        check(self._lib.sm_comm_init(self._h, buf))

    # -- uniform -----------------------------------------------------------------
    def write_uniform(self, uniform: SimSizeUniform) -> None:
        """queue.write_buffer(&sim_size_buffer, 0, bytes_of(&uniform)) -- main.rs:98."""
        check(self._lib.sm_set_params(self._h, C.byref(uniform)))

    def update_settings(self, settings: Settings) -> None:
        """update_settings(), main.rs:83-99."""
        self.settings = settings.clone()
        self.write_uniform(SimSizeUniform.new(self.width, self.height, settings.pheromone_decay_factor, settings))

    def read_uniform(self) -> SimSizeUniform:
        u = SimSizeUniform()
        check(self._lib.sm_get_params(self._h, C.byref(u)))
        return u

    # -- agents ------------------------------------------------------------------
    @property
    def agent_count(self) -> int:
        return int(self._lib.sm_agent_count(self._h))

    @property
    def local_agent_count(self) -> int:
        return int(self._lib.sm_local_agent_count(self._h))

    def init_agents(self, seed: int) -> None:
        """Seeded form of the start-up fill, main.rs:269-282 (done on the device)."""
        check(self._lib.sm_init_agents(self._h, C.c_uint64(seed)))

    def write_agents(self, agents: np.ndarray, first: int = 0) -> None:
        """queue.write_buffer(&agent_buffer, ..) -- main.rs:142, 994."""
        a = np.ascontiguousarray(agents, dtype=np.float32).reshape(-1, 4)
        check(self._lib.sm_upload_agents(self._h, _fptr(a), C.c_uint64(first), C.c_uint64(a.shape[0])))

    def read_agents(self, first: int = 0, n: Optional[int] = None, out: Optional[np.ndarray] = None):
        """copy_buffer_to_buffer + map_async read-back -- main.rs:121-131, 968-983."""
        n = self.agent_count - first if n is None else n
        if out is None:
            out = np.full((n, 4), np.nan, dtype=np.float32)
        owned = C.c_uint64(0)
        check(self._lib.sm_download_agents(self._h, _fptr(out), C.c_uint64(first), C.c_uint64(n), C.byref(owned)))
        self.last_owned = int(owned.value)
        return out

    def reassign_agent_speeds(self, seed: int) -> None:
        """reassign_agent_speeds(), main.rs:101-145 (no host round trip)."""
        check(self._lib.sm_reassign_speeds(self._h, C.c_uint64(seed)))

    def set_agent_count(self, n: int, seed: int) -> None:
        """N key: new buffer, fully re-randomised -- main.rs:682-791."""
        check(self._lib.sm_set_agent_count(self._h, C.c_uint64(n), C.c_uint64(seed)))
        self.settings.agent_count = int(n)

    # -- trail -------------------------------------------------------------------
    def clear_trail(self) -> None:
        """C key -- main.rs:909-913."""
        check(self._lib.sm_clear_trail(self._h))

    def write_trail(self, trail: np.ndarray, x0: int = 0, y0: int = 0) -> None:
        t = np.ascontiguousarray(trail, dtype=np.float32)
        h, w = t.shape
        check(self._lib.sm_upload_trail(self._h, _fptr(t), x0, y0, w, h, C.c_size_t(w)))

    def read_trail(self, x0: int = 0, y0: int = 0, w: Optional[int] = None, h: Optional[int] = None,
                   out: Optional[np.ndarray] = None) -> np.ndarray:
        w = self.width - x0 if w is None else w
        h = self.height - y0 if h is None else h
        if out is None:
            out = np.full((h, w), np.nan, dtype=np.float32)
        check(self._lib.sm_download_trail(self._h, _fptr(out), x0, y0, w, h, C.c_size_t(out.strides[0] // 4)))
        return out

    def trail_statistics(self) -> SmTrailStats:
        s = SmTrailStats()
        check(self._lib.sm_trail_statistics(self._h, C.byref(s)))
        return s

    def resize(self, width: int, height: int) -> None:
        """Window resize -- main.rs:954-1015 (agents rescaled, trail replaced by zeros)."""
        check(self._lib.sm_resize(self._h, width, height))
        self.width, self.height = int(width), int(height)

    # -- per-frame ---------------------------------------------------------------
    def step(self, n_steps: int = 1) -> None:
        """agents -> decay -> diffuse, one frame of main.rs:1163-1235 per step (asynchronous)."""
        check(self._lib.sm_step(self._h, n_steps))

    def diffuse_only(self, n_passes: int = 1) -> None:
        check(self._lib.sm_diffuse_only(self._h, n_passes))

    def sync(self) -> None:
        check(self._lib.sm_sync(self._h))

    # -- display pass (display.wgsl) and snapshots -----------------------------------
    def set_lut(self, lut) -> None:
        """`lut`: a LutData (lut_manager.rs) or 768 bytes laid out red ++ green ++ blue (main.rs:330-334)."""
        buf = lut.combined() if hasattr(lut, "combined") else np.ascontiguousarray(lut, dtype=np.uint8).reshape(-1)
        if buf.size != 768:
            raise ValueError("a LUT is 768 bytes (256 R, 256 G, 256 B)")
        check(self._lib.sm_set_lut(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint8))))

    def render(self, tex_width: int, tex_height: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """The display dispatch of main.rs:1202-1217: (tex_height, tex_width, 4) uint8 RGBA, letter-boxed."""
        if out is None:
            out = np.empty((tex_height, tex_width, 4), np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == (tex_height, tex_width, 4)
        check(self._lib.sm_render_rgba8(self._h, tex_width, tex_height, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def save_snapshot(self, path: str) -> None:
        check(self._lib.sm_save_snapshot(self._h, os.fsencode(path)))

    def load_snapshot(self, path: str) -> None:
        check(self._lib.sm_load_snapshot(self._h, os.fsencode(path)))

    # -- instrumentation ---------------------------------------------------------
    def set_timing_enabled(self, enabled: bool) -> None:
        check(self._lib.sm_set_timing_enabled(self._h, 1 if enabled else 0))

    def reset_timing(self) -> None:
        check(self._lib.sm_reset_timing(self._h))

    def timing(self) -> SmTiming:
        t = SmTiming()
        check(self._lib.sm_get_timing(self._h, C.byref(t)))
        return t

    @property
    def stream_handle(self) -> int:
        return int(self._lib.sm_stream(self._h) or 0)


def device_count() -> int:
    return int(_lib.load().sm_device_count())


def test_math(what: str, a, b=None, idx=None, device: int = 0):
    """Device evaluation of the arithmetic spec (tests only)."""
    lib = _lib.load()
    code = {"sincos": 0, "fmod": 1, "div9": 2, "hash01": 3}[what]
    a = np.ascontiguousarray(a, dtype=np.float32)
    o0 = np.empty_like(a)
    o1 = np.empty_like(a)
    bp = None
    ip = None
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.float32)
        bp = _fptr(b)
    if idx is not None:
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        ip = idx.ctypes.data_as(C.POINTER(C.c_int32))
    check(lib.sm_test_math(device, code, _fptr(a), bp, ip, _fptr(o0), _fptr(o1), C.c_uint64(a.size)))
    return (o0, o1) if what == "sincos" else o0
