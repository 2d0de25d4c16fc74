"""Runs the reference's shaders (their SOURCE TEXT, through tests/wgsl_interp.py) the way src/main.rs drives them.

TEST INFRASTRUCTURE.  /root/reference exists only in the build container: the golden vectors this module produces are
committed under tests/golden/wgsl_*.npz (tests/golden/make_wgsl_golden.py) together with the SHA-256 of the shader file
they came from; tests that re-run the shader skip when the reference tree is absent.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

from wgsl_interp import F32, Builtins, Interpreter, StorageArray, StorageTexture

REFERENCE_SRC = "/root/reference/src"

# SimSizeUniform, /root/reference/src/main.rs:29-46 -- member names as compute.wgsl:36-53 / display.wgsl:12-27 spell them
UNIFORM_FIELDS = ("width", "height", "decay_factor", "agent_jitter", "agent_speed_min", "agent_speed_max",
                  "agent_turn_speed", "agent_sensor_angle", "agent_sensor_distance", "diffusion_rate",
                  "pheromone_deposition_amount", "blur_radius", "blur_sigma")


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "compute.wgsl"))


def shader_source(name: str) -> str:
    with open(os.path.join(REFERENCE_SRC, name)) as f:
        return f.read()


def source_digest(text: str) -> str:
    return hashlib.sha256(text.encode()).hexdigest()


class SpecMath(Builtins):
    """sin / cos / float % of the arithmetic spec, evaluated by the oracle's C implementation:
    WGSL leaves their accuracy to the backend, the spec fixes it so that results can be compared bit for bit."""

    def __init__(self, oracle):
        self.so = oracle

    def sin(self, x):
        return F32(self.so.sincos(np.array([x], np.float32))[0][0])

    def cos(self, x):
        return F32(self.so.sincos(np.array([x], np.float32))[1][0])

    def fmod(self, a, b):
        return F32(self.so.fmod(np.array([a], np.float32), np.array([b], np.float32))[0])


def uniform_values(u) -> dict:
    """u: anything with SimSizeUniform's attributes (the ctypes mirror, the oracle's Params)."""
    return {k: getattr(u, k) for k in UNIFORM_FIELDS}


class ShaderSim:
    """compute.wgsl bound and dispatched like src/main.rs:1163-1235: `main` over the agents, `decay_trail` and
    `diffuse_trail` over the map, one dispatch each per frame, in that order."""

    def __init__(self, source, u, agents, trail, builtins=None):
        self.it = Interpreter(source, builtins)
        self.W, self.H = int(u.width), int(u.height)
        self.agents = StorageArray(np.ascontiguousarray(agents, np.float32).copy())
        self.trail = StorageArray(np.ascontiguousarray(trail, np.float32).reshape(-1).copy())
        self.it.bind("agents", self.agents)
        self.it.bind("trail_map", self.trail)
        vals = uniform_values(u)
        members = {m for m, _ in self.it.structs[self.it.gvars["sim_size"][3][0]]}
        self.it.bind_uniform("sim_size", **{k: v for k, v in vals.items() if k in members})

    def _agent_ids(self):
        # main.rs:1175-1180 dispatches ceil(N / 64) workgroups of 64 along x: the tail invocations return at
        # compute.wgsl:60-62 (run here too: they must not touch anything)
        n = self.agents.data.shape[0]
        wx = self.it.workgroup_size("main")[0]
        return [(i, 0, 0) for i in range((n + wx - 1) // wx * wx)]

    def _cell_ids(self, entry):
        wx, wy, _ = self.it.workgroup_size(entry)
        gx, gy = (self.W + wx - 1) // wx * wx, (self.H + wy - 1) // wy * wy      # whole workgroups: the guard at :152/:168 runs
        return [(x, y, 0) for y in range(gy) for x in range(gx)]

    def run_agents(self, schedule):
        self.it.dispatch("main", self._agent_ids(), schedule)

    def run_decay(self, schedule="sequential"):
        self.it.dispatch("decay_trail", self._cell_ids("decay_trail"), schedule)

    def run_diffuse(self, schedule):
        self.it.dispatch("diffuse_trail", self._cell_ids("diffuse_trail"), schedule)

    def frame(self, schedule, diffuse_schedule=None):
        self.run_agents(schedule)
        self.run_decay(schedule)
        self.run_diffuse(diffuse_schedule or schedule)

    @property
    def trail2d(self):
        return self.trail.data.reshape(self.H, self.W)


def run_display(source, u, trail, lut768, tex_w, tex_h):
    """display.wgsl:44-86 over a tex_w x tex_h storage texture; the LUT buffer is the 768 bytes widened to u32
    (main.rs:330-342)."""
    it = Interpreter(source)
    tex = StorageTexture(tex_w, tex_h)
    it.bind("trail_map", StorageArray(np.ascontiguousarray(trail, np.float32).reshape(-1).copy()))
    it.bind("display_tex", tex)
    it.bind("lut_data", StorageArray(np.asarray(lut768, np.uint8).astype(np.uint32)))
    it.bind_uniform("sim_size", **uniform_values(u))
    wx, wy, _ = it.workgroup_size("main")
    gx, gy = (tex_w + wx - 1) // wx * wx, (tex_h + wy - 1) // wy * wy
    it.dispatch("main", [(x, y, 0) for y in range(gy) for x in range(gx)])
    return tex.data
