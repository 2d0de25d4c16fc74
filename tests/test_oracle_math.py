"""The oracle's arithmetic spec against float64 / numpy ground truth (CPU)."""
import numpy as np

from conftest import bits_equal


def test_sincos_accuracy_small(oracle):
    x = np.random.default_rng(0).uniform(-30, 30, 1_000_000).astype(np.float32)
    s, c = oracle.sincos(x)
    xd = x.astype(np.float64)
    # WGSL demands 2^-11 absolute on [-pi, pi]; the spec'd polynomial is ~1 ulp
    assert np.abs(s - np.sin(xd)).max() < 1.3e-7
    assert np.abs(c - np.cos(xd)).max() < 1.3e-7


def test_sincos_accuracy_hash_range(oracle):
    # arguments of the jitter hash (compute.wgsl:117): 1e7 (1 M agents) .. 1.3e10 (1 B agents)
    x = np.random.default_rng(1).uniform(-1.4e10, 1.4e10, 1_000_000).astype(np.float32)
    s, c = oracle.sincos(x)
    xd = x.astype(np.float64)
    assert np.abs(s - np.sin(xd)).max() < 1.3e-7
    assert np.abs(c - np.cos(xd)).max() < 1.3e-7


def test_sincos_path_boundary_and_specials(oracle):
    x = np.array([0.0, -0.0, 8192.0, -8192.0, np.nextafter(np.float32(8192), np.float32(9000)), 1e15, 3e38,
                  np.inf, -np.inf, np.nan], dtype=np.float32)
    s, c = oracle.sincos(x)
    assert s[0] == 0 and not np.signbit(s[0]) and c[0] == 1
    assert s[1] == 0 and c[1] == 1      # the spec does not preserve the sign of zero
    xd = x[2:5].astype(np.float64)
    assert np.abs(s[2:5] - np.sin(xd)).max() < 1.3e-7
    assert np.all(np.isfinite(s[5:7])) and np.all(np.abs(s[5:7]) <= 1.0)
    assert np.all(np.isnan(s[7:])) and np.all(np.isnan(c[7:]))


def test_fmod_is_ieee(oracle):
    rng = np.random.default_rng(2)
    a = np.concatenate([rng.uniform(-100, 100, 200000), rng.uniform(-1e9, 1e9, 200000)]).astype(np.float32)
    for b in (np.float32(6.2831855), np.float32(1920), np.float32(1080)):
        bb = np.full_like(a, b)
        assert bits_equal(oracle.fmod(a, bb), np.fmod(a, bb))


def test_div9(oracle):
    a = np.random.default_rng(3).uniform(0, 9, 500000).astype(np.float32)
    assert bits_equal(oracle.div9(a), a / np.float32(9.0))


def test_hash_in_unit_interval_and_deterministic(oracle):
    rng = np.random.default_rng(4)
    n = 500000
    idx = rng.integers(0, 2**31 - 1, n).astype(np.int32)
    x = rng.uniform(0, 4096, n).astype(np.float32)
    y = rng.uniform(0, 4096, n).astype(np.float32)
    r = oracle.hash01(idx, x, y)
    assert r.min() >= 0.0 and r.max() <= 1.0
    assert 0.49 < r.mean() < 0.51
    assert bits_equal(r, oracle.hash01(idx, x, y))
    # definition: fract(sin(f32(idx)*12.9898 + x*78.233 + y*37.719) * 43758.5453), f32, no contraction
    arg = (idx.astype(np.float32) * np.float32(12.9898) + x * np.float32(78.233)) + y * np.float32(37.719)
    s, _ = oracle.sincos(arg)
    v = s * np.float32(43758.5453)
    assert bits_equal(r, v - np.floor(v))


def test_rng_distribution(oracle):
    W, H = 1920, 1080
    a = oracle.init_agents(400000, W, H, 30.0, 50.0, seed=1)
    assert a[:, 0].min() >= 0 and a[:, 0].max() <= W
    assert a[:, 1].min() >= 0 and a[:, 1].max() <= H
    assert a[:, 2].min() >= 0 and a[:, 2].max() < 2 * np.pi + 1e-6
    assert a[:, 3].min() >= 30 and a[:, 3].max() <= 50
    assert abs(a[:, 0].mean() / W - 0.5) < 0.01 and abs(a[:, 1].mean() / H - 0.5) < 0.01
    # 24-bit mantissa convention of rand::random::<f32>()
    u = a[:, 0] / np.float32(W)
    b = oracle.init_agents(400000, W, H, 30.0, 50.0, seed=2)
    assert not np.array_equal(a, b)
    # chunked generation == one-shot generation (counter based)
    c = oracle.init_agents(1000, W, H, 30.0, 50.0, seed=1, first_id=5000)
    assert bits_equal(c, a[5000:6000])
    assert u.max() < 1.0 + 1e-6
