"""ctypes binding of the CPU oracle (oracle/libslime_oracle.so).

TEST INFRASTRUCTURE -- the checker, never the thing shipped or measured as the
product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  Parity pin (see the header of
slime_oracle.c): the reference has no golden vectors and cannot run in this image;
this restatement of /root/reference/src/compute.wgsl is pinned bit for bit to the
shader's own source text executed by tests/wgsl_interp.py (tests/golden/wgsl_*.npz,
tests/test_wgsl_reference.py), with sin / cos / float % fixed by the arithmetic spec.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libslime_oracle.so")


class Params(C.Structure):
    """SimSizeUniform, /root/reference/src/main.rs:29-46 (56 bytes)."""

    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("decay_factor", C.c_float),
        ("agent_jitter", C.c_float),
        ("agent_speed_min", C.c_float),
        ("agent_speed_max", C.c_float),
        ("agent_turn_speed", C.c_float),
        ("agent_sensor_angle", C.c_float),
        ("agent_sensor_distance", C.c_float),
        ("diffusion_rate", C.c_float),
        ("pheromone_deposition_amount", C.c_float),
        ("blur_radius", C.c_float),
        ("blur_sigma", C.c_float),
        ("pad", C.c_uint32),
    ]


_FLAVOUR = "gcc -O2 -march=x86-64-v3"


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc, -ffp-contract=off, OpenMP).  With SM_ORACLE_NATIVE=1 in the
    environment (bench.py's CPU arm) a second library is built on this very machine at -O3 -march=native and loaded
    instead; if that build fails the portable one is used."""
    global _SO, _FLAVOUR, _lib
    srcs = [os.path.join(_HERE, f) for f in ("slime_oracle.c", "sm_oracle_math.h", "Makefile")]
    src_m = max(os.path.getmtime(f) for f in srcs)
    portable = os.path.join(_HERE, "libslime_oracle.so")
    if force or not os.path.exists(portable) or os.path.getmtime(portable) < src_m:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    if os.environ.get("SM_ORACLE_NATIVE") == "1" and _lib is None:
        native = os.path.join(_HERE, "libslime_oracle_native.so")
        stamp = native + ".host"
        host = _host_id()
        fresh = os.path.exists(native) and os.path.getmtime(native) >= src_m and os.path.exists(stamp) and open(stamp).read() == host
        if not fresh:
            r = subprocess.run(["make", "-C", _HERE, "-s", "-B", "native"], capture_output=True, text=True)
            fresh = r.returncode == 0 and os.path.exists(native)
            if fresh:
                with open(stamp, "w") as f:
                    f.write(host)
        if fresh:
            _SO, _FLAVOUR = native, "gcc -O3 -march=native (built on this host)"
    return _SO


def _host_id() -> str:
    """-march=native code must not travel to another CPU: the native build is tagged with the CPU it was made on."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build_flavour() -> str:
    return _FLAVOUR


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.so_params_size.restype = C.c_int
        assert _lib.so_params_size() == 56 == C.sizeof(Params)
        _lib.so_max_threads.restype = C.c_int
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def make_params(width, height, **kw) -> Params:
    """Defaults = Settings::default(), /root/reference/src/settings.rs:8-27."""
    p = Params(
        width=width, height=height, decay_factor=10.0, agent_jitter=0.0,
        agent_speed_min=30.0, agent_speed_max=50.0, agent_turn_speed=0.43,
        agent_sensor_angle=0.3, agent_sensor_distance=20.0, diffusion_rate=1.0,
        pheromone_deposition_amount=1.0, blur_radius=2.0, blur_sigma=1.0, pad=0,
    )
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def max_threads() -> int:
    return lib().so_max_threads()


def set_threads(n: int) -> None:
    lib().so_set_threads(C.c_int(n))


# ---- arithmetic spec --------------------------------------------------------
def sincos(x):
    x = _f32(x)
    s = np.empty_like(x)
    c = np.empty_like(x)
    lib().so_sincos_array(_ptr(x, C.c_float), _ptr(s, C.c_float), _ptr(c, C.c_float), C.c_uint64(x.size))
    return s, c


def fmod(a, b):
    a = _f32(a)
    b = _f32(b)
    r = np.empty_like(a)
    lib().so_fmod_array(_ptr(a, C.c_float), _ptr(b, C.c_float), _ptr(r, C.c_float), C.c_uint64(a.size))
    return r


def div9(a):
    a = _f32(a)
    r = np.empty_like(a)
    lib().so_div9_array(_ptr(a, C.c_float), _ptr(r, C.c_float), C.c_uint64(a.size))
    return r


def hash01(idx, x, y):
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    x = _f32(x)
    y = _f32(y)
    r = np.empty_like(x)
    lib().so_hash01_array(_ptr(idx, C.c_int32), _ptr(x, C.c_float), _ptr(y, C.c_float), _ptr(r, C.c_float), C.c_uint64(x.size))
    return r


# ---- state ------------------------------------------------------------------
def init_agents(n, W, H, speed_min, speed_max, seed, first_id=0):
    a = np.empty((n, 4), dtype=np.float32)
    lib().so_init_agents(_ptr(a, C.c_float), C.c_uint64(first_id), C.c_uint64(n), C.c_uint32(W), C.c_uint32(H),
                         C.c_float(speed_min), C.c_float(speed_max), C.c_uint64(seed))
    return a


def reassign_speeds(agents, speed_min, speed_max, seed, first_id=0):
    assert agents.dtype == np.float32 and agents.flags.c_contiguous
    lib().so_reassign_speeds(_ptr(agents, C.c_float), C.c_uint64(first_id), C.c_uint64(agents.shape[0]),
                             C.c_float(speed_min), C.c_float(speed_max), C.c_uint64(seed))
    return agents


def rescale_agents(agents, oldW, oldH, newW, newH):
    assert agents.dtype == np.float32 and agents.flags.c_contiguous
    lib().so_rescale_agents(_ptr(agents, C.c_float), C.c_uint64(agents.shape[0]), C.c_uint32(oldW), C.c_uint32(oldH),
                            C.c_uint32(newW), C.c_uint32(newH))
    return agents


# ---- passes -----------------------------------------------------------------
def agents_phase_split(agents, trail, counts, p: Params, ids=None):
    """In place on `agents` (n,4) f32 and `counts` (H,W) u32; `trail` read only."""
    assert agents.dtype == np.float32 and agents.flags.c_contiguous
    assert trail.dtype == np.float32 and trail.flags.c_contiguous
    assert counts.dtype == np.uint32 and counts.flags.c_contiguous
    idp = None
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        idp = _ptr(ids, C.c_uint32)
    lib().so_agents_phase_split(_ptr(agents, C.c_float), idp, C.c_uint64(agents.shape[0]), _ptr(trail, C.c_float),
                                _ptr(counts, C.c_uint32), C.byref(p))


def agents_sequential(agents, trail, p: Params):
    assert agents.dtype == np.float32 and agents.flags.c_contiguous
    assert trail.dtype == np.float32 and trail.flags.c_contiguous
    lib().so_agents_sequential(_ptr(agents, C.c_float), C.c_uint64(agents.shape[0]), _ptr(trail, C.c_float), C.byref(p))


def deposit_merge(trail, counts, dep):
    lib().so_deposit_merge(_ptr(trail, C.c_float), _ptr(counts, C.c_uint32), C.c_uint64(trail.size), C.c_float(dep))


def decay(trail, decay_factor):
    assert trail.dtype == np.float32 and trail.flags.c_contiguous
    lib().so_decay(_ptr(trail, C.c_float), C.c_uint64(trail.size), C.c_float(decay_factor))


def diffuse(trail, diffusion_rate):
    """Jacobi 3x3 toroidal mean; returns a new array."""
    assert trail.dtype == np.float32 and trail.flags.c_contiguous
    H, W = trail.shape
    out = np.empty_like(trail)
    lib().so_diffuse(_ptr(trail, C.c_float), _ptr(out, C.c_float), C.c_uint32(W), C.c_uint32(H), C.c_float(diffusion_rate))
    return out


def diffuse_inplace_raster(trail, diffusion_rate):
    H, W = trail.shape
    lib().so_diffuse_inplace_raster(_ptr(trail, C.c_float), C.c_uint32(W), C.c_uint32(H), C.c_float(diffusion_rate))


def trail_pass(trail, p: Params, counts=None, gauss_radius=0, gauss_sigma=0.0):
    """The engine's fused pass (merge -> decay -> blur); returns the new field, clears counts."""
    assert trail.dtype == np.float32 and trail.flags.c_contiguous
    H, W = trail.shape
    assert (W, H) == (p.width, p.height)
    out = np.empty_like(trail)
    scratch = np.empty_like(trail)
    cp = _ptr(counts, C.c_uint32) if counts is not None else None
    if gauss_radius > 0:
        scratch2 = np.empty_like(trail)
        lib().so_trail_pass_gauss(_ptr(trail, C.c_float), cp, _ptr(out, C.c_float), _ptr(scratch, C.c_float),
                                  _ptr(scratch2, C.c_float), C.c_uint32(W), C.c_uint32(H), C.byref(p),
                                  C.c_int(gauss_radius), C.c_float(gauss_sigma))
    else:
        lib().so_trail_pass(_ptr(trail, C.c_float), cp, _ptr(out, C.c_float), _ptr(scratch, C.c_float),
                            C.c_uint32(W), C.c_uint32(H), C.byref(p))
    return out


def display(trail, lut768, tex_w, tex_h):
    """display.wgsl:29-86: trail -> letter-boxed RGBA8 (tex_h, tex_w, 4) through a 768-byte planar LUT."""
    assert trail.dtype == np.float32 and trail.flags.c_contiguous
    H, W = trail.shape
    lut = np.ascontiguousarray(lut768, dtype=np.uint8)
    assert lut.size == 768
    out = np.empty((tex_h, tex_w, 4), np.uint8)
    lib().so_display(_ptr(trail, C.c_float), C.c_uint32(W), C.c_uint32(H), _ptr(lut, C.c_uint8),
                     _ptr(out, C.c_uint8), C.c_uint32(tex_w), C.c_uint32(tex_h))
    return out


def gauss_weights(R, sigma):
    w = np.empty(2 * R + 1, dtype=np.float32)
    lib().so_gauss_weights(_ptr(w, C.c_float), C.c_int(R), C.c_float(sigma))
    return w


class Sim:
    """Whole-simulation driver over the oracle passes (pass order of main.rs:1163-1235)."""

    def __init__(self, p: Params, agents, trail=None, ids=None):
        self.p = p
        self.agents = np.ascontiguousarray(agents, dtype=np.float32).copy()
        H, W = p.height, p.width
        self.trail = np.zeros((H, W), np.float32) if trail is None else np.ascontiguousarray(trail, np.float32).copy()
        self.counts = np.zeros((H, W), np.uint32)
        self._tmp = np.empty((H, W), np.float32)
        self._scr = np.empty((H, W), np.float32)
        self.ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)

    def step(self, n=1):
        idp = _ptr(self.ids, C.c_uint32) if self.ids is not None else None
        lib().so_step_phase_split(_ptr(self.agents, C.c_float), idp, C.c_uint64(self.agents.shape[0]),
                                  _ptr(self.trail, C.c_float), _ptr(self.counts, C.c_uint32),
                                  _ptr(self._tmp, C.c_float), _ptr(self._scr, C.c_float), C.byref(self.p), C.c_int(n))

    def step_sequential(self, n=1, inplace_diffuse=False):
        lib().so_step_sequential(_ptr(self.agents, C.c_float), C.c_uint64(self.agents.shape[0]),
                                 _ptr(self.trail, C.c_float), _ptr(self._tmp, C.c_float), C.byref(self.p),
                                 C.c_int(n), C.c_int(1 if inplace_diffuse else 0))


def reference_dispatch_hits(n):
    hits = np.zeros(n, dtype=np.uint8)
    lib().so_reference_dispatch_hits(_ptr(hits, C.c_uint8), C.c_uint64(n))
    return hits
