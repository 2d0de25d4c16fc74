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


def edge_agents(W, H):
    """Agents that exercise the rare paths: zero / negative-zero headings (dead-hash test), huge and
    negative headings (slow sincos path, general fmod), positions outside the map, on the seams,
    non-finite state."""
    f = np.float32
    rows = []
    for ang in (0.0, -0.0, 1e-30, 6.2831855, -6.2831855, 12.566371, -3.0, 100.0, 5000.0, 1e6, -1e6, 3e9, 1e20, np.inf, np.nan):
        rows.append([W * 0.5, H * 0.5, ang, 40.0])
        rows.append([W * 0.25 + 0.5, H * 0.75 + 0.25, ang, 35.0])
    for x, y in ((0.0, 0.0), (-0.0, -0.0), (W - 1e-3, H - 1e-3), (float(W), float(H)), (-5.0, -7.0), (2.5 * W, 3.5 * H),
                 (-1e7, 1e7), (1e31, 1.0), (1.0, 1e31), (np.inf, 1.0), (1.0, np.nan), (W - 0.5, 0.25), (0.25, H - 0.5)):
        for ang in (0.0, 1.0, 3.0, 4.5):
            rows.append([x, y, ang, 45.0])
    rows.append([10.0, 10.0, 1.0, np.nan])
    rows.append([10.0, 10.0, 1.0, -5.0])
    rows.append([10.0, 10.0, 1.0, 1e9])
    return np.array(rows, dtype=f)
