"""The oracle must keep reproducing the committed golden vectors (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import bits_equal
from presets_util import PRESET_NAMES, preset_uniform, to_oracle_params

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, "golden_" + name.lower().replace(" ", "_") + ".npz"))


@pytest.mark.parametrize("name", PRESET_NAMES)
def test_oracle_reproduces_golden(oracle, name):
    g = load(name)
    H, W = g["trail1"].shape
    u = preset_uniform(name, W, H)
    assert bytes(u) == g["params"].tobytes()          # the preset mirror packs the same 56 bytes
    p = to_oracle_params(oracle, u)
    ag0 = oracle.init_agents(g["agents0"].shape[0], W, H, u.agent_speed_min, u.agent_speed_max, int(g["seed"]))
    assert bits_equal(ag0, g["agents0"])
    sim = oracle.Sim(p, ag0)
    sim.step(1)
    assert bits_equal(sim.agents, g["agents1"]) and bits_equal(sim.trail, g["trail1"])
    sim.step(24)
    assert bits_equal(sim.agents, g["agents25"]) and bits_equal(sim.trail, g["trail25"])
    assert bits_equal(oracle.trail_pass(g["field"], p, counts=None), g["diffused"])
