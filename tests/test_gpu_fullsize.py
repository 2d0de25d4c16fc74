"""Full-size checks at BASELINE.json configs[1] (16,777,216 agents, 4096x4096): two steps bit-exact
against the oracle, then size-independent properties over a longer run.  -m gpu."""
import numpy as np
import pytest

import slime_mold_b200 as sm
from conftest import bits_equal, mismatch_report
from presets_util import preset_uniform, to_oracle_params

pytestmark = pytest.mark.gpu

N, W, H = 16_777_216, 4096, 4096


def test_config2_two_steps_bit_exact(oracle, engine_lib):
    s = sm.init_preset_manager().get_preset("Firecracker Trees").settings   # jitter on: hash args reach 2e8
    be = sm.CudaBackend.new(W, H, s, agent_count=N)
    be.init_agents(seed=1)
    ref = oracle.init_agents(N, W, H, s.agent_speed_min, s.agent_speed_max, 1)
    sim = oracle.Sim(to_oracle_params(oracle, preset_uniform("Firecracker Trees", W, H)), ref)
    sim.step(2)
    be.step(2)
    a, t = be.read_agents(), be.read_trail()
    assert bits_equal(a, sim.agents), mismatch_report(a, sim.agents, "agents")
    assert bits_equal(t, sim.trail), mismatch_report(t, sim.trail, "trail")
    be.close()


def test_config2_properties_long_run(engine_lib):
    s = sm.init_preset_manager().get_preset("Default").settings
    runs = []
    for flags in (0, sm.SM_FLAG_NO_SORT):
        be = sm.CudaBackend.new(W, H, s, agent_count=N, flags=flags)
        be.init_agents(seed=3)
        be.step(40)
        a, t = be.read_agents(), be.read_trail()
        st = be.trail_statistics()
        be.close()
        # invariants of the step loop
        assert np.isfinite(a).all()
        assert a[:, 0].min() >= 0 and a[:, 0].max() <= W and a[:, 1].min() >= 0 and a[:, 1].max() <= H
        assert a[:, 2].min() >= 0 and a[:, 2].max() < 6.2831856
        assert a[:, 3].min() >= s.agent_speed_min and a[:, 3].max() <= s.agent_speed_max
        assert t.min() >= 0.0 and t.max() <= 1.0
        assert abs(st.sum - t.sum(dtype=np.float64)) < 1e-7 * t.size and st.max == t.max()
        runs.append((a, t))
    # the cell sort only permutes storage: results are identical with and without it (order-free deposits)
    assert bits_equal(runs[0][0], runs[1][0]) and bits_equal(runs[0][1], runs[1][1])


def test_diffusion_mass_and_fixed_point_8192(engine_lib):
    Wb = Hb = 8192
    s = sm.Settings.default().clone(pheromone_decay_factor=0.0)
    be = sm.CudaBackend.new(Wb, Hb, s, agent_count=1)
    rng = np.random.default_rng(0)
    field = rng.random((Hb, Wb), dtype=np.float32)
    be.write_trail(field)
    be.diffuse_only(4)
    st = be.trail_statistics()
    # decay 0: the 3x3 mean conserves mass up to rounding (relative 1e-6)
    assert abs(st.sum - field.sum(dtype=np.float64)) < 1e-6 * field.size
    # constant field 0.5 is an exact fixed point (0.5*9 and /9 exact)
    be.write_trail(np.full((Hb, Wb), 0.5, np.float32))
    be.diffuse_only(3)
    out = be.read_trail()
    assert (out == np.float32(0.5)).all()
    be.close()
