"""Cross-check of the oracle's logic against an independent numpy transliteration of
compute.wgsl that uses numpy's own sin/cos (stand-in for 'some other wgpu backend')."""
import numpy as np
import pytest

import numpy_wgsl as nw
from conftest import bits_equal
from presets_util import PRESET_NAMES, preset_uniform, random_trail, to_oracle_params

# north-star tolerance for one step against a different sin/cos: sensor positions move by <= sd * 2^-11,
# positions by <= 1e-3 px unless a sensor comparison flips (then the heading differs by the turn speed)
POS_TOL = 2e-3


@pytest.mark.parametrize("name", PRESET_NAMES)
def test_one_step_matches_numpy_within_tolerance(oracle, name):
    W, H, N = 384, 256, 30000
    u = preset_uniform(name, W, H)
    p = to_oracle_params(oracle, u)
    trail = random_trail(W, H, seed=11)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, seed=4)
    a_or = ag.copy(); cnt = np.zeros((H, W), np.uint32)
    oracle.agents_phase_split(a_or, trail, cnt, p)
    a_np, cnt_np = nw.agents_pass(ag, trail, u)
    exact = np.all(a_or.view(np.uint32) == a_np.view(np.uint32), axis=1).mean()
    # speed is arithmetic-free (clamp): always identical
    assert bits_equal(a_or[:, 3], a_np[:, 3])
    if u.agent_jitter == 0.0:
        # no hash influence: headings identical except where a sensor comparison flipped on a 1-ulp sin/cos
        # difference; positions within tolerance for everyone whose heading agrees
        same_heading = a_or[:, 2] == a_np[:, 2]
        assert same_heading.mean() > 0.995
        d = np.abs(a_or[same_heading, :2] - a_np[same_heading, :2])
        d = np.minimum(d, np.array([W, H], np.float32) - d)
        assert d.max() < POS_TOL
        assert exact > 0.90
    else:
        # the hash multiplies a 1-ulp difference of sin() by 43758.5453: the jitter draw differs by up to
        # ~2 * 43758 * 2^-23 = 0.0105 (times the jitter strength) between two correct sin implementations,
        # and wraps by a full unit when fract() crosses an integer (no two backends agree here)
        tol = 1e-5 + 0.011 * u.agent_jitter
        dh = np.abs(a_or[:, 2] - a_np[:, 2])
        dh = np.minimum(dh, 2 * np.pi - dh)
        same_heading = dh < tol
        assert same_heading.mean() > 0.97
        d = np.abs(a_or[same_heading, :2] - a_np[same_heading, :2])
        d = np.minimum(d, np.array([W, H], np.float32) - d)
        assert d.max() < POS_TOL + u.agent_speed_max * 0.016 * tol
    # deposits land on the same cells for agents that agree
    assert abs(int(cnt.sum()) - int(cnt_np.sum())) <= 2


@pytest.mark.parametrize("name", ["Default", "Curls", "Threads"])
@pytest.mark.parametrize("shape", [(64, 48), (37, 23), (5, 3), (2, 2), (1, 1)])
def test_trail_pass_bit_exact_vs_numpy(oracle, name, shape):
    W, H = shape
    u = preset_uniform(name, W, H)
    p = to_oracle_params(oracle, u)
    rng = np.random.default_rng(W * 100 + H)
    trail = rng.random((H, W), dtype=np.float32)
    counts = (rng.random((H, W)) < 0.3).astype(np.uint32) * rng.integers(1, 5, (H, W)).astype(np.uint32)
    out = oracle.trail_pass(trail, p, counts=counts.copy())
    ref = nw.trail_pass(trail, counts, u)
    assert bits_equal(out, ref)
    out2 = oracle.trail_pass(trail, p, counts=None)
    assert bits_equal(out2, nw.trail_pass(trail, None, u))


def test_fractional_deposit_trail_pass(oracle):
    W, H = 40, 30
    u = preset_uniform("Default", W, H)
    u.pheromone_deposition_amount = 0.15
    p = to_oracle_params(oracle, u)
    rng = np.random.default_rng(5)
    trail = rng.random((H, W), dtype=np.float32)
    counts = rng.integers(0, 9, (H, W)).astype(np.uint32)
    assert bits_equal(oracle.trail_pass(trail, p, counts=counts.copy()), nw.trail_pass(trail, counts, u))
