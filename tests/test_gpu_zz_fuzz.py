"""Randomised parity on the GPU: the cases of tests/test_wgsl_fuzz.py (tiny maps, parameters far outside the presets,
agents outside the map) through the C ABI against the oracle, bit for bit.  -m gpu.  (The oracle itself is pinned to the
reference's shader source on these very cases, tests/test_wgsl_fuzz.py.)"""
import numpy as np
import pytest

import slime_mold_b200 as sm
from conftest import bits_equal, mismatch_report
from presets_util import to_oracle_params
from test_wgsl_fuzz import describe, random_case

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("seed", [10, 11, 12])
def test_engine_equals_oracle_on_random_cases(oracle, engine_lib, seed):
    rng = np.random.default_rng(seed)
    for case in range(30):
        u, ag, tr = random_case(rng)
        sim = oracle.Sim(to_oracle_params(oracle, u), ag, tr)
        with sm.CudaBackend.new(int(u.width), int(u.height), sm.Settings.default(), agent_count=ag.shape[0], device=0) as be:
            be.write_uniform(u)
            be.write_agents(ag)
            be.write_trail(tr)
            for k in range(3):
                sim.step(1)
                be.step(1)
                a, t = be.read_agents(), be.read_trail()
                assert bits_equal(a, sim.agents), (case, k, describe(u), mismatch_report(a, sim.agents, "agents"))
                assert bits_equal(t, sim.trail), (case, k, describe(u), mismatch_report(t, sim.trail, "trail"))
