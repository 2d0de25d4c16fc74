"""Helpers shared by the tests: preset -> oracle parameter block, seeded states."""
import numpy as np

from slime_mold_b200 import SimSizeUniform, init_preset_manager

PRESET_NAMES = ["Default", "Sponge", "Firecracker Trees", "Threads", "Curls", "Waves", "Snake", "Mesh"]


def preset_uniform(name: str, width: int, height: int) -> SimSizeUniform:
    s = init_preset_manager().get_preset(name).settings
    return SimSizeUniform.new(width, height, s.pheromone_decay_factor, s)


def to_oracle_params(so, u: SimSizeUniform):
    p = so.Params()
    for f, _ in SimSizeUniform._fields_:
        setattr(p, "pad" if f == "_pad" else f, getattr(u, f))
    return p


def random_trail(W, H, seed, density=0.3):
    rng = np.random.default_rng(seed)
    t = rng.random((H, W), dtype=np.float32)
    t[rng.random((H, W)) > density] = 0.0
    return t
