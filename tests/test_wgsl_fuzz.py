"""Randomised parity, off the beaten track: tiny maps (1..11 x 1..9), a handful of agents, parameters far outside the
presets -- negative decay / jitter / deposit / sensor distance, turn speeds above TAU, speeds below zero, diffusion rates
outside [0, 1] -- and agents that start outside the map or with headings of several hundred radians.

  * where /root/reference exists (the build container): the reference's SHADER SOURCE (tests/wgsl_interp.py) against the
    oracle, bit for bit, in the sequential schedule (two frames) and in the lockstep schedule (one dispatch);
  * everywhere: the engine's __host__ __device__ arithmetic (tests/hostcheck) against the oracle on the same cases.
"""
import ctypes as C

import numpy as np
import pytest

import wgsl_reference as wr
from conftest import bits_equal, mismatch_report
from presets_util import preset_uniform, to_oracle_params


def random_case(rng):
    W, H, N = int(rng.integers(1, 12)), int(rng.integers(1, 10)), int(rng.integers(1, 25))
    u = preset_uniform("Default", W, H)

    def pick(*c):
        return float(c[rng.integers(0, len(c))])
    u.decay_factor = pick(0, 10, 100, -5, 1500, rng.random() * 50)
    u.agent_jitter = pick(0, 0, 0.1, 1.0, 5.0, -0.3)
    u.agent_speed_min = pick(0, 10, 30, -20, 100)
    u.agent_speed_max = pick(50, 5, 200, 1000, u.agent_speed_min)
    u.agent_turn_speed = pick(0.43, 0, 7.0, -1.0, 0.02)
    u.agent_sensor_angle = pick(0.3, 0, -0.5, 3.2, 1.34)
    u.agent_sensor_distance = pick(1.0, 2.5, 0, 20, -3, 0.4)
    u.diffusion_rate = pick(1.0, 0, 0.5, -0.5, 2.0)
    u.pheromone_deposition_amount = pick(1.0, 0.3, 2.5, 0, -0.5, 1.0)
    ag = np.stack([rng.random(N) * W * rng.choice([1, 1, 1, 3, -1]), rng.random(N) * H * rng.choice([1, 1, 1, -2, 2]),
                   (rng.random(N) * 20 - 5) * rng.choice([1, 1, 100]), rng.random(N) * 100 - 10], axis=1).astype(np.float32)
    tr = (rng.random((H, W)) * rng.choice([1, 1, 1.5]) - rng.choice([0, 0, 0.2])).astype(np.float32)
    return u, ag, tr


def describe(u):
    return {f: getattr(u, f) for f in wr.UNIFORM_FIELDS}


@pytest.mark.skipif(not wr.have_reference(), reason="/root/reference is not present on this machine")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_shader_source_equals_oracle_on_random_cases(oracle, seed):
    src = wr.shader_source("compute.wgsl")
    rng = np.random.default_rng(seed)
    for case in range(14):
        u, ag, tr = random_case(rng)
        p = to_oracle_params(oracle, u)
        sh = wr.ShaderSim(src, u, ag, tr, wr.SpecMath(oracle))
        sim = oracle.Sim(p, ag, tr)
        for k in range(2):
            sh.frame("sequential")
            sim.step_sequential(1, inplace_diffuse=True)
            assert bits_equal(sh.agents.data, sim.agents), (case, describe(u), mismatch_report(sh.agents.data, sim.agents, "agents"))
            assert bits_equal(sh.trail2d, sim.trail), (case, describe(u), mismatch_report(sh.trail2d, sim.trail, "trail"))
        sh = wr.ShaderSim(src, u, ag, tr, wr.SpecMath(oracle))
        sim = oracle.Sim(p, ag, tr)
        sh.run_agents("lockstep")
        oracle.agents_phase_split(sim.agents, sim.trail, sim.counts, p)
        assert bits_equal(sh.agents.data, sim.agents), (case, describe(u), mismatch_report(sh.agents.data, sim.agents, "agents"))
        if u.pheromone_deposition_amount >= 1.0 and tr.min() >= 0:       # where lockstep and phase_split deposits coincide
            oracle.deposit_merge(sim.trail, sim.counts, p.pheromone_deposition_amount)
            assert bits_equal(sh.trail2d, sim.trail), (case, describe(u), mismatch_report(sh.trail2d, sim.trail, "trail"))


@pytest.mark.parametrize("seed", [10, 11, 12, 13])
def test_engine_arithmetic_on_host_equals_oracle_on_random_cases(oracle, hostcheck, seed):
    def P(a, t):
        return a.ctypes.data_as(C.POINTER(t))
    rng = np.random.default_rng(seed)
    for case in range(40):
        u, ag, tr = random_case(rng)
        p = to_oracle_params(oracle, u)
        H, W = tr.shape
        sim = oracle.Sim(p, ag, tr)
        a, t = ag.copy(), tr.copy()
        cn = np.zeros((H, W), np.uint32)
        out = np.empty_like(t)
        for k in range(3):
            sim.step(1)
            hostcheck.hc_agents_phase_split(P(a, C.c_float), None, C.c_uint64(a.shape[0]), P(t, C.c_float), P(cn, C.c_uint32), C.byref(p))
            hostcheck.hc_trail_pass(P(t, C.c_float), P(cn, C.c_uint32), P(out, C.c_float), C.byref(p))
            t, out = out, t
            assert bits_equal(a, sim.agents), (case, k, describe(u), mismatch_report(a, sim.agents, "agents"))
            assert bits_equal(t, sim.trail), (case, k, describe(u), mismatch_report(t, sim.trail, "trail"))


@pytest.mark.skipif(not wr.have_reference(), reason="/root/reference is not present on this machine")
def test_display_shader_source_equals_oracle_on_random_cases(oracle):
    """display.wgsl (interpreted) against the oracle's display restatement: random map and texture shapes (both letter-box
    orientations, magnified and minified), random LUTs, cells at the quantisation edges (k/255 +- half a step), NaN / inf."""
    src = wr.shader_source("display.wgsl")
    rng = np.random.default_rng(7)
    for case in range(24):
        W, H = int(rng.integers(1, 40)), int(rng.integers(1, 30))
        tw, th = int(rng.integers(1, 48)), int(rng.integers(1, 40))
        u = preset_uniform("Default", W, H)
        tr = (rng.random((H, W)) * 1.6 - 0.3).astype(np.float32)
        for _ in range(3):
            tr[rng.integers(0, H), rng.integers(0, W)] = rng.choice([np.nan, np.inf, -np.inf, 1.0, 0.0, 0.99999994, 1.0000001,
                                                                      0.5 / 255, 254.5 / 255, 255.5 / 255])
        lut = rng.integers(0, 256, 768, dtype=np.uint8)
        got = oracle.display(tr, lut, tw, th)
        ref = wr.run_display(src, u, tr, lut, tw, th)
        assert np.array_equal(got, ref), (case, W, H, tw, th, int((got != ref).sum()))
