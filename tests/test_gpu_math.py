"""Device evaluation of the arithmetic spec (sm_test_math through the C ABI) against the
oracle, bit for bit.  -m gpu."""
import numpy as np
import pytest

from conftest import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu


def specials():
    return np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 8192.0, -8192.0, 8192.001, 1e15, 2e15, 3e38, 1e-40,
                     6.2831855, -6.2831855, 12.566371], dtype=np.float32)


def test_sincos_device_bits(oracle, engine_lib):
    from slime_mold_b200.backend import test_math
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.integers(0, 2**32, 4_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32),
        rng.uniform(-10, 10, 2_000_000).astype(np.float32), rng.uniform(-9000, 9000, 1_000_000).astype(np.float32),
        rng.uniform(-2e10, 2e10, 2_000_000).astype(np.float32), specials()])
    s0, c0 = oracle.sincos(x)
    s1, c1 = test_math("sincos", x)
    assert bits_equal(s0, s1), mismatch_report(s0, s1, "sin")
    assert bits_equal(c0, c1), mismatch_report(c0, c1, "cos")


@pytest.mark.parametrize("b", [6.2831855, 1920.0, 1080.0, 4096.0, 32768.0, 65536.0, 1.0, 7.0, 37.0])
def test_fmod_device_bits(oracle, engine_lib, b):
    from slime_mold_b200.backend import test_math
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.uniform(-3 * b, 3 * b, 1_000_000), rng.uniform(-1e7, 1e7, 500_000),
                        rng.uniform(-1e12, 1e12, 100_000)]).astype(np.float32)
    a = np.concatenate([a, rng.integers(0, 2**32, 1_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32), specials(),
                        np.float32(b) * np.arange(-5, 6, dtype=np.float32)])
    bb = np.full_like(a, np.float32(b))
    r0 = oracle.fmod(a, bb)
    r1 = test_math("fmod", a, bb)
    assert bits_equal(r0, r1), mismatch_report(r0, r1, "fmod")


def test_div9_device_bits(oracle, engine_lib):
    from slime_mold_b200.backend import test_math
    rng = np.random.default_rng(2)
    d = np.concatenate([rng.integers(0, 2**32, 4_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32),
                        rng.uniform(0, 9, 2_000_000).astype(np.float32), specials()])
    d = d[~((d == 0) & np.signbit(d))]
    r0 = oracle.div9(d)
    r1 = test_math("div9", d)
    assert bits_equal(r0, r1), mismatch_report(r0, r1, "div9")


def test_hash_device_bits(oracle, engine_lib):
    from slime_mold_b200.backend import test_math
    rng = np.random.default_rng(3)
    n = 4_000_000
    idx = rng.integers(0, 2**31 - 1, n).astype(np.int32)
    x = rng.uniform(0, 32768, n).astype(np.float32)
    y = rng.uniform(0, 32768, n).astype(np.float32)
    r0 = oracle.hash01(idx, x, y)
    r1 = test_math("hash01", x, y, idx)
    assert bits_equal(r0, r1), mismatch_report(r0, r1, "hash01")
