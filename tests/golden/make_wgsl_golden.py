"""Generates tests/golden/wgsl_*.npz: outputs of the REFERENCE'S OWN SHADER SOURCE, executed here.

    python tests/golden/make_wgsl_golden.py          (needs /root/reference; about two minutes)

The reference cannot be built in this image (Rust + wgpu; no rustc, no Vulkan), but its hot path is two text files,
/root/reference/src/compute.wgsl and display.wgsl.  tests/wgsl_interp.py interprets that text (it contains no knowledge
of what the shaders compute); tests/wgsl_reference.py binds the buffers and issues the dispatches in the order of
src/main.rs:1163-1235.  The vectors written here are what pins the C oracle -- and, on the GPU box, the CUDA engine
-- to the reference:

  wgsl_<case>.npz   params (the 56 uniform bytes), agents0, trail0, and per frame k = 1..FRAMES
                      seq_agents<k>, seq_trail<k>    schedule "sequential" (invocation i finishes before i+1 starts,
                                                     storage live: agents in index order, Gauss-Seidel diffuse)
                      lock_agents<k>, lock_trail<k>  schedule "lockstep" (every load sees the buffers as they were when
                                                     the dispatch started) -- the engine's phase_split semantics
                    sin / cos / float % are those of the arithmetic spec (WGSL leaves their
                    accuracy to the backend), so the comparison with the oracle and the engine is bit for bit.
                    libm_agents1: the first lockstep frame again with numpy's libm sin / cos, for the tolerance test.
  wgsl_edge.npz     one `main` dispatch over the edge-case agents of tests/presets_util.py (NaN / inf / huge headings,
                    positions outside the map and on the seams), both schedules
  wgsl_display.npz  display.wgsl over four texture shapes (letter-box left/right and top/bottom, exact fit, tiny) with
                    NaN / inf / out-of-range cells in the field

Every file records the SHA-256 of the shader text it was produced from; tests/test_wgsl_reference.py re-runs a slice of
each case when /root/reference is present and checks the digest.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import slime_oracle as so  # noqa: E402  (only for SpecMath: sin / cos / fmod of the arithmetic spec)
from presets_util import edge_agents, preset_uniform, random_trail  # noqa: E402
from wgsl_reference import ShaderSim, SpecMath, run_display, shader_source, source_digest  # noqa: E402

W, H, N, FRAMES, SEED = 48, 40, 400, 3, 11

# (file tag, preset, overrides of the uniform): presets with long sensors are shortened to fit the small map; "lowdep"
# has dep < 1 (sequential schedule only: lockstep and phase_split differ there by design)
CASES = [
    ("default", "Default", {}),
    ("sponge", "Sponge", {}),
    ("waves", "Waves", {}),
    ("snake", "Snake", {"agent_sensor_distance": 11.0}),
    ("mesh_jitter", "Mesh", {"agent_jitter": 0.35, "diffusion_rate": 0.6}),
    ("lowdep", "Default", {"pheromone_deposition_amount": 0.3, "agent_jitter": 0.1, "decay_factor": 35.0}),
]

# tiny and ragged maps (the engine's generic kernels: W % 4 != 0, fewer cells than a warp, a single cell): short sensors so
# that some land inside, agents wrap around the map several times
SMALL = {"agent_sensor_distance": 2.5, "agent_jitter": 0.15}
SMALL_CASES = [
    ("tiny_1x1", 1, 1, 6, SMALL),
    ("tiny_2x3", 2, 3, 12, SMALL),
    ("ragged_13x7", 13, 7, 60, SMALL),
    ("ragged_37x5", 37, 5, 80, {"agent_sensor_distance": 1.5, "diffusion_rate": 0.5}),
]
SMALL_FRAMES = 4


def case_uniform(preset, over, w=W, h=H):
    u = preset_uniform(preset, w, h)
    for k, v in over.items():
        setattr(u, k, v)
    return u


def initial_state(u, seed, n=N):
    w, h = int(u.width), int(u.height)
    rng = np.random.default_rng(seed)
    ag = np.stack([rng.random(n) * w, rng.random(n) * h, rng.random(n) * 6.2831855,
                   u.agent_speed_min + rng.random(n) * (u.agent_speed_max - u.agent_speed_min)], axis=1).astype(np.float32)
    return ag, random_trail(w, h, seed=seed + 1, density=0.35 if w * h > 100 else 0.8)


def make_case(src, tag, preset, over, w=W, h=H, n=N, frames=FRAMES):
    u = case_uniform(preset, over, w, h)
    ag0, tr0 = initial_state(u, SEED, n)
    out = dict(params=np.frombuffer(bytes(u), dtype=np.uint8), agents0=ag0, trail0=tr0, frames=frames,
               shader_sha256=source_digest(src))
    for sched, key in (("sequential", "seq"), ("lockstep", "lock")):
        sim = ShaderSim(src, u, ag0, tr0, SpecMath(so))
        for k in range(1, frames + 1):
            sim.frame(sched)
            out[f"{key}_agents{k}"] = sim.agents.data.copy()
            out[f"{key}_trail{k}"] = sim.trail2d.copy()
    sim = ShaderSim(src, u, ag0, tr0)                      # numpy's libm
    sim.run_agents("lockstep")
    out["libm_agents1"] = sim.agents.data.copy()
    np.savez_compressed(os.path.join(HERE, f"wgsl_{tag}.npz"), **out)


def make_edge(src):
    u = case_uniform("Sponge", {"agent_jitter": 0.2})
    ag0 = edge_agents(W, H)
    tr0 = random_trail(W, H, seed=5, density=0.5)
    out = dict(params=np.frombuffer(bytes(u), dtype=np.uint8), agents0=ag0, trail0=tr0, shader_sha256=source_digest(src))
    for sched, key in (("sequential", "seq"), ("lockstep", "lock")):
        sim = ShaderSim(src, u, ag0, tr0, SpecMath(so))
        sim.run_agents(sched)
        out[f"{key}_agents1"] = sim.agents.data.copy()
        out[f"{key}_trail1"] = sim.trail2d.copy()
    np.savez_compressed(os.path.join(HERE, "wgsl_edge.npz"), **out)


DISPLAY_SHAPES = [(64, 36), (30, 50), (48, 40), (17, 13)]


def display_inputs():
    u = case_uniform("Default", {})
    rng = np.random.default_rng(21)
    lut = rng.integers(0, 256, 768, dtype=np.uint8)
    tr = random_trail(W, H, seed=9, density=0.6)
    tr[0, 0], tr[1, 1], tr[2, 2], tr[3, 3], tr[4, 4] = np.nan, 2.0, -1.0, np.inf, -np.inf
    return u, lut, tr


def make_display(src):
    u, lut, tr = display_inputs()
    out = dict(params=np.frombuffer(bytes(u), dtype=np.uint8), lut=lut, trail=tr, shader_sha256=source_digest(src),
               shapes=np.array(DISPLAY_SHAPES))
    for tw, th in DISPLAY_SHAPES:
        out[f"rgba_{tw}x{th}"] = run_display(src, u, tr, lut, tw, th)
    np.savez_compressed(os.path.join(HERE, "wgsl_display.npz"), **out)


if __name__ == "__main__":
    so.build()
    csrc, dsrc = shader_source("compute.wgsl"), shader_source("display.wgsl")
    only_small = "--small" in sys.argv          # the small cases were added later: same generator, same shader files
    for tag, preset, over in ([] if only_small else CASES):
        t = time.time()
        make_case(csrc, tag, preset, over)
        print(f"wgsl_{tag}.npz  {time.time() - t:.1f} s")
    for tag, w, h, n, over in SMALL_CASES:
        make_case(csrc, tag, "Default", over, w, h, n, SMALL_FRAMES)
        print(f"wgsl_{tag}.npz")
    if not only_small:
        make_edge(csrc)
        make_display(dsrc)
        print("wgsl_edge.npz, wgsl_display.npz")
