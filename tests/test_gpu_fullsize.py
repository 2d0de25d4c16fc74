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


# ---- maps beyond 2^31 cells: the IdxT = int64_t instantiations, and the LDG sampler fallback above the texture-gather limit ----
def _host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2**30
    except Exception:
        return 0.0


def test_step_on_a_map_beyond_2_31_cells(oracle, engine_lib):
    """46344 x 46400 = 2.15e9 cells: cell offsets need 64 bits from row 46341 on, and the map is wider than the 32768-texel
    gather limit, so the agent kernel samples the row-major field (k_agents<.., int64_t, FetchLinear<int64_t>>).  200,000 agents
    placed where the offsets exceed 2^31 (and across the toroidal seam), three full steps, against the oracle on the same map:
    agents bit for bit, the trail rows they can have touched bit for bit, the field's sum."""
    if _host_ram_gb() < 80:
        pytest.skip("the full-size oracle needs ~45 GB of host memory")
    Wb, Hb, Nb, steps = 46344, 46400, 200_000, 3
    s = sm.init_preset_manager().get_preset("Waves").settings            # jitter + turning: every branch is live
    u = sm.SimSizeUniform.new(Wb, Hb, s.pheromone_decay_factor, s)
    p = to_oracle_params(oracle, u)
    rng = np.random.default_rng(5)
    ag = np.empty((Nb, 4), np.float32)
    ag[:, 0] = rng.uniform(38000, Wb, Nb)
    ag[:, 1] = rng.uniform(46250, Hb, Nb)
    ag[:5000, 1] = rng.uniform(0, 25, 5000)                               # the other side of the seam
    ag[:, 2] = rng.uniform(0, 6.28, Nb)
    ag[:, 3] = rng.uniform(s.agent_speed_min, s.agent_speed_max, Nb)
    with sm.CudaBackend.new(Wb, Hb, s, agent_count=Nb) as be:
        be.write_agents(ag)
        be.step(steps)
        a = be.read_agents()
        bottom = be.read_trail(y0=46200, h=200)
        top = be.read_trail(y0=0, h=64)
        st = be.trail_statistics()
    trail = np.zeros((Hb, Wb), np.float32)
    counts = np.zeros((Hb, Wb), np.uint32)
    ref = ag.copy()
    for _ in range(steps):
        oracle.agents_phase_split(ref, trail, counts, p)
        trail = oracle.trail_pass(trail, p, counts=counts)
    assert bits_equal(a, ref), mismatch_report(a, ref, "agents")
    assert bits_equal(bottom, trail[46200:46400]), mismatch_report(bottom, trail[46200:46400], "trail, rows 46200..46399")
    assert bits_equal(top, trail[:64]), mismatch_report(top, trail[:64], "trail, rows 0..63")
    assert not trail[64:46200].any()                                       # nothing else can have been touched
    assert abs(st.sum - trail.sum(dtype=np.float64)) < 1e-6 * max(1.0, st.sum)


def test_diffusion_only_65536_squared(oracle, engine_lib):
    """BASELINE config 5's largest map: 65536^2 cells, 16 GiB per field.  The blur is translation invariant, so two patches --
    one in the far corner, where offsets exceed 2^32 -- must evolve exactly like the same patches on a small toroidal map
    under the oracle; with decay 0 the field's mass is conserved."""
    S = 65536
    s = sm.Settings.default().clone(pheromone_decay_factor=0.0, pheromone_diffusion_rate=0.7)
    rng = np.random.default_rng(11)
    patch = rng.random((24, 40), dtype=np.float32)
    passes = 5
    small = np.zeros((64, 96), np.float32)
    small[20:44, 28:68] = patch
    ps = to_oracle_params(oracle, sm.SimSizeUniform.new(96, 64, s.pheromone_decay_factor, s))
    for _ in range(passes):
        small = oracle.trail_pass(small, ps)
    with sm.CudaBackend.new(S, S, s, agent_count=1) as be:
        for (x0, y0) in ((65400, 65480), (1000, 33000)):
            be.write_trail(patch, x0=x0, y0=y0)
        be.diffuse_only(passes)
        for (x0, y0) in ((65400, 65480), (1000, 33000)):
            got = be.read_trail(x0=x0 - 28, y0=y0 - 20, w=96, h=64)
            assert bits_equal(got, small), mismatch_report(got, small, f"patch at ({x0}, {y0})")
        st = be.trail_statistics()
    assert abs(st.sum - 2.0 * float(patch.sum(dtype=np.float64))) < 2e-3
    assert int(st.nonzero) == 2 * int(np.count_nonzero(small))
