"""Hand-derived known answers for the oracle, read directly off
/root/reference/src/compute.wgsl (line numbers in the comments).  These pin the
oracle's *logic*; bit-level arithmetic is defined by the oracle itself
(the reference ships no vectors; the pin to its shader source is tests/test_wgsl_reference.py)."""
import numpy as np
import pytest

from conftest import bits_equal

TAU32 = np.float32(6.28318530718)


def one_step(so, p, agents, trail=None):
    H, W = p.height, p.width
    a = np.ascontiguousarray(agents, dtype=np.float32).copy()
    t = np.zeros((H, W), np.float32) if trail is None else trail.astype(np.float32).copy()
    c = np.zeros((H, W), np.uint32)
    so.agents_phase_split(a, t, c, p)
    return a, c


def test_zero_field_straight_motion(oracle):
    # L = R = C = 0 -> "do nothing" branch (:110-112): heading unchanged, straight move (:125-127)
    p = oracle.make_params(256, 128)
    ag = np.array([[100.0, 50.0, 0.0, 40.0], [10.5, 20.25, 1.0, 35.0], [200.0, 100.0, 4.0, 50.0]], np.float32)
    a, c = one_step(oracle, p, ag)
    for i in range(3):
        d = np.float64(np.float32(ag[i, 3]) * np.float32(0.016))
        assert a[i, 2] == ag[i, 2]
        assert abs(a[i, 0] - (ag[i, 0] + d * np.cos(np.float64(ag[i, 2])))) < 1e-4
        assert abs(a[i, 1] - (ag[i, 1] + d * np.sin(np.float64(ag[i, 2])))) < 1e-4
        assert c[int(a[i, 1]), int(a[i, 0])] == 1            # :136-140
    assert c.sum() == 3


def test_speed_is_clamped_and_stored(oracle):
    p = oracle.make_params(64, 64, agent_speed_min=30.0, agent_speed_max=50.0)   # :72, :144
    ag = np.array([[10, 10, 0, 5.0], [10, 20, 0, 500.0], [10, 30, 0, 42.0]], np.float32)
    a, _ = one_step(oracle, p, ag)
    assert list(a[:, 3]) == [30.0, 50.0, 42.0]
    assert abs(a[0, 0] - (10 + 30 * 0.016)) < 1e-5 and abs(a[1, 0] - (10 + 50 * 0.016)) < 1e-5


def test_steering_branches(oracle):
    W = H = 128
    p = oracle.make_params(W, H, agent_turn_speed=0.43, agent_sensor_angle=0.5, agent_sensor_distance=20.0)
    x, y, ang = 64.0, 64.0, 0.0
    def probe(angle):  # sensor position for a heading
        return x + 20 * np.cos(angle), y + 20 * np.sin(angle)
    def blob(t, pos, v):
        px, py = int(np.floor(pos[0])), int(np.floor(pos[1]))
        t[py - 1:py + 3, px - 1:px + 3] = v
    ag = np.array([[x, y, ang, 40.0]], np.float32)
    # centre strongest -> keep (:98)
    t = np.zeros((H, W), np.float32); blob(t, probe(0.0), 1.0); blob(t, probe(-0.5), 0.5); blob(t, probe(0.5), 0.2)
    a, _ = one_step(oracle, p, ag, t); assert a[0, 2] == 0.0
    # left (angle - sa) strongest -> angle -= turn, wrapped into [0, 2pi) (:100-104, :121-122)
    t = np.zeros((H, W), np.float32); blob(t, probe(-0.5), 1.0)
    a, _ = one_step(oracle, p, ag, t)
    assert abs(a[0, 2] - (2 * np.pi - 0.43)) < 1e-6
    # right strongest -> angle += turn (:105-109)
    t = np.zeros((H, W), np.float32); blob(t, probe(0.5), 1.0)
    a, _ = one_step(oracle, p, ag, t)
    assert a[0, 2] == np.float32(0.0) + np.float32(0.43)
    # L == R > C -> neither `>` holds -> keep (:110-112)
    t = np.zeros((H, W), np.float32); blob(t, probe(-0.5), 0.7); blob(t, probe(0.5), 0.7)
    a, _ = one_step(oracle, p, ag, t); assert a[0, 2] == 0.0
    # C == L > R: first test fails (strict >), L > R -> turn left
    t = np.zeros((H, W), np.float32); blob(t, probe(0.0), 0.7); blob(t, probe(-0.5), 0.7)
    a, _ = one_step(oracle, p, ag, t); assert abs(a[0, 2] - (2 * np.pi - 0.43)) < 1e-6


def test_turn_uses_min_of_turn_speed_and_tau(oracle):
    # :102-104  diff = (angle - TAU) - angle ~ -TAU; min(turn_speed, |diff|) caps the turn at ~TAU
    W = H = 128
    p = oracle.make_params(W, H, agent_turn_speed=100.0, agent_sensor_angle=0.5)
    ag = np.array([[64, 64, 1.0, 40.0]], np.float32)
    t = np.zeros((H, W), np.float32)
    lx, ly = 64 + 20 * np.cos(0.5), 64 + 20 * np.sin(0.5)
    t[int(ly) - 1:int(ly) + 3, int(lx) - 1:int(lx) + 3] = 1.0
    a, _ = one_step(oracle, p, ag, t)
    diff = np.float32(np.float32(1.0) - TAU32) - np.float32(1.0)
    expect = np.float32(1.0) + np.float32(min(np.float32(100.0), abs(diff))) * np.float32(-1.0)
    expect = np.fmod(expect, TAU32)
    if expect < 0:
        expect = expect + TAU32
    assert a[0, 2] == expect


def test_sensing_is_not_toroidal_and_last_row_col_read_zero(oracle):
    # :14-16  x0 < 0 || x1 >= W || y0 < 0 || y1 >= H -> 0.0
    W, H = 64, 32
    p = oracle.make_params(W, H, agent_sensor_angle=0.5, agent_sensor_distance=5.0, agent_turn_speed=0.25)
    t = np.ones((H, W), np.float32)
    # heading +x near the right edge: C and both side sensors land at x >= W-1 -> all 0 -> keep heading
    a, _ = one_step(oracle, p, np.array([[W - 3.0, 16.0, 0.0, 40.0]], np.float32), t)
    assert a[0, 2] == 0.0
    # same position heading -x: all sensors inside a constant field -> L == R == C -> keep
    a, _ = one_step(oracle, p, np.array([[W - 3.0, 16.0, np.pi, 40.0]], np.float32), t)
    assert a[0, 2] == np.float32(np.pi)
    # heading +x, 6.5 px from the edge: centre tap x0 = 62 = W-2 is the LAST valid column; side taps inside
    # C = L = R = 1 -> keep.  One pixel further right the centre reads 0 -> L == R -> still keep;
    # make L weaker than R by zeroing the upper half -> turns right (+)
    t2 = t.copy(); t2[:16, :] = 0.0
    a, _ = one_step(oracle, p, np.array([[W - 6.5, 16.5, 0.0, 40.0]], np.float32), t2)
    assert a[0, 2] == np.float32(0.25)


def test_toroidal_position_wrap_and_deposit(oracle):
    # :130-133 wrap, :136-140 deposit at the wrapped cell
    W, H = 64, 32
    p = oracle.make_params(W, H, agent_speed_min=50.0, agent_speed_max=50.0)
    ag = np.array([[W - 0.25, 5.0, 0.0, 50.0], [0.25, 5.0, np.pi, 50.0], [10.0, H - 0.25, np.pi / 2, 50.0]], np.float32)
    a, c = one_step(oracle, p, ag)
    assert 0.5 < a[0, 0] < 0.6 and W - 0.6 < a[1, 0] < W - 0.5 and 0.5 < a[2, 1] < 0.6
    assert c[5, 0] == 1 and c[4, W - 1] + c[5, W - 1] == 1 and c[0, 10] + c[0, 9] == 1


def test_jitter_uses_premove_xy_and_index(oracle):
    # :115-118  angle += (fract(sin(idx*12.9898 + x*78.233 + y*37.719)*43758.5453)*2 - 1) * jitter
    W = H = 128
    p = oracle.make_params(W, H, agent_jitter=0.5)
    ag = np.array([[33.25, 77.5, 1.0, 40.0]] * 3, np.float32)
    a, _ = one_step(oracle, p, ag)
    r = oracle.hash01(np.arange(3, dtype=np.int32), ag[:, 0], ag[:, 1])
    expect = ag[:, 2] + (r * np.float32(2.0) - np.float32(1.0)) * np.float32(0.5)
    assert bits_equal(a[:, 2], expect)
    assert len(set(a[:, 2].tolist())) == 3        # the index enters the hash
    # persistent ids: the hash follows the id array, not the storage position
    ids = np.array([2, 0, 1], np.uint32)
    b = ag.copy(); cnt = np.zeros((H, W), np.uint32)
    oracle.agents_phase_split(b, np.zeros((H, W), np.float32), cnt, p, ids=ids)
    assert bits_equal(b[:, 2], a[ids, 2])


def test_decay_known_answer(oracle):
    # :159-160  t = max(t - decay_factor*0.001, 0)
    t = np.array([[1.0, 0.005, 0.0, 0.5]], np.float32)
    oracle.decay(t, 10.0)
    d = np.float32(10.0) * np.float32(0.001)
    assert bits_equal(t, np.maximum(np.array([[1.0, 0.005, 0.0, 0.5]], np.float32) - d, np.float32(0)))
    assert t[0, 1] == 0.0 and t[0, 2] == 0.0


def test_diffuse_known_answers(oracle):
    # :176-194  toroidal 3x3 mean, mix(t, mean, rate)
    t = np.zeros((8, 8), np.float32); t[0, 0] = 1.0
    o = oracle.diffuse(t, 1.0)
    ninth = np.float32(1.0) / np.float32(9.0)
    for (y, x) in [(0, 0), (0, 1), (1, 0), (1, 1), (7, 7), (7, 0), (0, 7), (1, 7), (7, 1)]:
        assert o[y, x] == ninth
    assert np.count_nonzero(o) == 9
    o = oracle.diffuse(t, 0.0)
    assert bits_equal(o, t)
    o = oracle.diffuse(t, 0.25)   # mix = t*(1-r) + mean*r
    assert o[0, 0] == np.float32(1.0) * np.float32(0.75) + ninth * np.float32(0.25)
    # rate is clamped to [0, 1] (:173)
    assert bits_equal(oracle.diffuse(t, 7.0), oracle.diffuse(t, 1.0))
    assert bits_equal(oracle.diffuse(t, -3.0), oracle.diffuse(t, 0.0))
    # constant field is a fixed point (0.5 * 9 and /9 are exact)
    c = np.full((5, 7), 0.5, np.float32)
    assert bits_equal(oracle.diffuse(c, 1.0), c)


def test_diffuse_mass_conservation(oracle):
    t = np.random.default_rng(0).random((64, 96), dtype=np.float32)
    o = oracle.diffuse(t, 1.0)
    assert abs(o.sum(dtype=np.float64) - t.sum(dtype=np.float64)) < 1e-3 * t.size * 1e-3


def test_tiny_maps(oracle):
    # (x+dx+W)%W with W < 3: neighbours alias (:183-184)
    t = np.array([[0.9]], np.float32)
    o = oracle.diffuse(t, 1.0)
    s = np.float32(0)
    for _ in range(9):
        s = s + np.float32(0.9)
    assert o[0, 0] == s / np.float32(9)
    t = np.array([[1.0, 0.0]], np.float32)
    o = oracle.diffuse(t, 1.0)
    # row sums for x=0: (x-1)=1,(x)=0,(x+1)=1 -> 0+1+0 per row, three rows -> 3/9
    assert o[0, 0] == np.float32(3) / np.float32(9) and o[0, 1] == np.float32(6) / np.float32(9)


def test_deposit_merge_equals_sequential_rmw_when_dep_ge_1(oracle):
    # compute.wgsl:140  trail = clamp(trail + dep, 0, 1); with dep >= 1 any k >= 1 gives exactly 1.0
    W = H = 64
    p = oracle.make_params(W, H, agent_sensor_distance=1e6)    # every sensor tap is outside -> no steering
    ag = oracle.init_agents(20000, W, H, 30, 50, seed=5)
    t0 = np.random.default_rng(1).random((H, W), dtype=np.float32)
    a1 = ag.copy(); cnt = np.zeros((H, W), np.uint32)
    oracle.agents_phase_split(a1, t0, cnt, p)
    t1 = t0.copy(); oracle.deposit_merge(t1, cnt, 1.0)
    a2 = ag.copy(); t2 = t0.copy()
    oracle.agents_sequential(a2, t2, p)
    assert bits_equal(a1, a2) and bits_equal(t1, t2)
    assert cnt.sum() == 0   # merge clears the counts


def test_deposit_merge_fractional(oracle):
    t = np.array([[0.2, 0.2, 0.95, 0.0]], np.float32)
    c = np.array([[0, 3, 1, 40]], np.uint32)
    oracle.deposit_merge(t, c, 0.1)
    exp = np.array([0.2, min(1.0, np.float32(0.2) + np.float32(3) * np.float32(0.1)), 1.0, 1.0], np.float32)
    assert bits_equal(t[0], exp)


def test_pass_order_agents_decay_diffuse(oracle):
    # main.rs:1163-1235: agents -> decay -> diffuse.  One agent on a zero field, dep 1, decay 10, rate 1.
    W = H = 16
    p = oracle.make_params(W, H)
    sim = oracle.Sim(p, np.array([[8.2, 8.2, 0.0, 40.0]], np.float32))
    sim.step(1)
    cx, cy = int(sim.agents[0, 0]), int(sim.agents[0, 1])
    v = (np.float32(1.0) - np.float32(10.0) * np.float32(0.001))
    s = np.float32(0)
    for k in range(9):
        s = s + (v if k == 4 else np.float32(0))
    assert sim.trail[cy, cx] == s / np.float32(9)
    assert np.count_nonzero(sim.trail) == 9


def test_reference_dispatch_quirk_table(oracle):
    # main.rs:1175-1179 dispatches (min(wg,65535), ceil(wg/x)) groups of 64, the shader
    # linearises with stride 65535 (compute.wgsl:60).  The engine uses the linear map instead.
    for n, reached in [(1_000_000, 1_000_000), (4_194_240, 4_194_240), (10_000_000, 4_325_310)]:
        hits = oracle.reference_dispatch_hits(n)
        assert int((hits > 0).sum()) == reached
    assert oracle.reference_dispatch_hits(10_000_000).max() == 3


def test_rescale_and_reassign(oracle):
    a = oracle.init_agents(1000, 640, 480, 30, 50, seed=3)
    b = a.copy()
    oracle.rescale_agents(b, 640, 480, 800, 600)          # main.rs:985-989
    fx = np.float32(800) / np.float32(640); fy = np.float32(600) / np.float32(480)
    assert bits_equal(b[:, 0], a[:, 0] * fx) and bits_equal(b[:, 1], a[:, 1] * fy)
    c = a.copy()
    oracle.reassign_speeds(c, 70, 80, seed=9)             # main.rs:133-137
    assert c[:, 3].min() >= 70 and c[:, 3].max() <= 80 and bits_equal(c[:, :3], a[:, :3])
