"""Display pass and snapshots on the GPU (through the C ABI), bit for bit against the oracle."""
import numpy as np
import pytest

import slime_mold_b200 as sm
from conftest import bits_equal, mismatch_report
from presets_util import preset_uniform, random_trail, to_oracle_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,tw,th", [(192, 108, 192, 108), (192, 108, 320, 200), (100, 37, 63, 200), (64, 64, 1, 1),
                                      (33, 77, 500, 40), (1920, 1080, 1600, 900), (512, 512, 2048, 1024), (256, 128, 30, 17)])
def test_display_equals_oracle(oracle, engine_lib, W, H, tw, th):
    rng = np.random.default_rng(W + 7 * th)
    t = (rng.random((H, W), dtype=np.float32) * np.float32(1.4) - np.float32(0.2)).astype(np.float32)
    t[rng.integers(0, H, 5), rng.integers(0, W, 5)] = np.nan
    t[0, 0] = np.inf
    t[H - 1, W - 1] = -np.inf
    lut = rng.integers(0, 256, 768).astype(np.uint8)
    with sm.CudaBackend.new(W, H, sm.Settings.default(), agent_count=16, device=0) as be:
        be.write_trail(t)
        with pytest.raises(sm.SlimeError):
            be.render(tw, th)                                   # no LUT yet
        be.set_lut(lut)
        frame = be.render(tw, th)
    ref = oracle.display(t, lut, tw, th)
    assert np.array_equal(frame, ref), f"{np.count_nonzero(frame != ref)} bytes differ"


def test_display_after_steps_and_lut_data(oracle, engine_lib, tmp_path):
    W, H, N, steps = 320, 180, 40_000, 25
    s = sm.init_preset_manager().get_preset("Default").settings
    lm = sm.LutManager()
    lut = lm.load_lut("gray_r")
    with sm.CudaBackend.new(W, H, s, agent_count=N, device=0) as be:
        be.init_agents(seed=5)
        be.step(steps)
        be.set_lut(lut)                                         # a LutData, as main.rs:156-166 loads it
        frame = be.render(400, 300)
        trail = be.read_trail()
        be.diffuse_only(1)                                      # no agent pass: the frame is the field as it stands
        frame_after = be.render(400, 300)
        trail_after = be.read_trail()
    # the reference draws between its decay and diffuse dispatches (main.rs:1184-1217): the field of the last step after
    # the deposits and the decay, before the blur
    u = preset_uniform("Default", W, H)
    p = to_oracle_params(oracle, u)
    sim = oracle.Sim(p, oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 5))
    sim.step(steps - 1)
    oracle.agents_phase_split(sim.agents, sim.trail, sim.counts, p)
    pre = sim.trail.copy()
    oracle.deposit_merge(pre, sim.counts, u.pheromone_deposition_amount)
    oracle.decay(pre, u.decay_factor)
    assert bits_equal(trail, oracle.diffuse(pre, u.diffusion_rate))                # same step, same state
    assert np.array_equal(frame, oracle.display(pre, lut.combined(), 400, 300))
    assert not np.array_equal(frame, oracle.display(trail, lut.combined(), 400, 300))
    assert np.array_equal(frame_after, oracle.display(trail_after, lut.combined(), 400, 300))
    assert len(np.unique(frame[..., 0])) > 8 and (frame[..., 3] == 255).all()      # something was drawn
    sm.write_png(str(tmp_path / "frame.png"), frame)


def test_display_pre_diffuse_frame_with_fractional_deposits(oracle, engine_lib):
    """u32 deposit counts (deposition amount < 1): the frame recomputes clamp(t + k*dep, 0, 1) and the decay per texel."""
    W, H, N, steps = 256, 144, 60_000, 7
    s = sm.init_preset_manager().get_preset("Waves").settings.clone(pheromone_deposition_amount=0.3)
    lut = np.random.default_rng(3).integers(0, 256, 768).astype(np.uint8)
    u = sm.SimSizeUniform.new(W, H, s.pheromone_decay_factor, s)
    p = to_oracle_params(oracle, u)
    with sm.CudaBackend.new(W, H, s, agent_count=N, device=0) as be:
        be.init_agents(seed=2)
        be.step(steps)
        be.set_lut(lut)
        frame = be.render(300, 200)
    sim = oracle.Sim(p, oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 2))
    sim.step(steps - 1)
    oracle.agents_phase_split(sim.agents, sim.trail, sim.counts, p)
    pre = sim.trail.copy()
    oracle.deposit_merge(pre, sim.counts, u.pheromone_deposition_amount)
    oracle.decay(pre, u.decay_factor)
    assert np.array_equal(frame, oracle.display(pre, lut, 300, 200))


@pytest.mark.parametrize("preset", ["Default", "Waves"])
def test_snapshot_restore_continues_bit_exactly(oracle, engine_lib, tmp_path, preset):
    W, H, N, s1, s2 = 384, 256, 120_000, 21, 19
    s = sm.init_preset_manager().get_preset(preset).settings
    path = str(tmp_path / "state.smb")
    with sm.CudaBackend.new(W, H, s, agent_count=N, device=0) as be:
        be.init_agents(seed=9)
        be.step(s1)                                             # crosses a cell sort: the slot order is permuted
        be.save_snapshot(path)
        be.step(s2)
        a_cont, t_cont = be.read_agents(), be.read_trail()
    other = sm.init_preset_manager().get_preset("Sponge").settings          # overwritten by the snapshot's parameter block
    with sm.CudaBackend.new(W, H, other, agent_count=N, device=0) as be:
        be.load_snapshot(path)
        be.step(s2)
        a_rest, t_rest = be.read_agents(), be.read_trail()
    assert bits_equal(a_cont, a_rest), mismatch_report(a_rest, a_cont, "agents")
    assert bits_equal(t_cont, t_rest), mismatch_report(t_rest, t_cont, "trail")
    u = preset_uniform(preset, W, H)
    sim = oracle.Sim(to_oracle_params(oracle, u), oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 9))
    sim.step(s1 + s2)
    assert bits_equal(a_rest, sim.agents) and bits_equal(t_rest, sim.trail)
    # a snapshot of another geometry is refused
    with sm.CudaBackend.new(W, H // 2, s, agent_count=N, device=0) as be:
        with pytest.raises(sm.SlimeError):
            be.load_snapshot(path)
    with sm.CudaBackend.new(W, H, s, agent_count=N, device=0) as be:
        with pytest.raises(sm.SlimeError):
            be.load_snapshot(str(tmp_path / "missing.smb"))
