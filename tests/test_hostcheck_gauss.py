"""The streaming Gaussian kernel (slime_mold_b200/csrc/gauss_stream.cuh) on the CPU: tests/hostcheck runs the kernel
body -- the very source the GPU compiles -- one host thread per CUDA thread with a pthread barrier per CTA, and the
result is compared bit for bit with the oracle's definition of the extension.  Covers every radius, all three deposit
representations, ragged last tiles / chunks, the toroidal seam on all four sides and chunk heights that are not a
multiple of the batch.  (The device instantiation is checked by tests/test_gpu_parity.py, -m gpu.)"""
import ctypes as C

import numpy as np
import pytest

import slime_mold_b200 as sm
from conftest import bits_equal, mismatch_report
from presets_util import random_trail, to_oracle_params


def P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


GUARD = 4      # rows of guard band around every array the kernel writes: a store outside the map must show up


def guarded(shape, dtype, fill):
    H, W = shape
    full = np.full((H + 2 * GUARD, W), fill, dtype)
    return full, full[GUARD:GUARD + H]


def guards_intact(full, fill):
    g = np.concatenate([full[:GUARD].ravel(), full[-GUARD:].ravel()])
    return bool(np.all(np.isnan(g))) if isinstance(fill, float) and np.isnan(fill) else bool(np.all(g == fill))


def run_stream(hostcheck, oracle, field, p, R, sigma, chunk, cm=0, counts=None, want_surf=False, kernel="stream"):
    H, W = field.shape
    # the input sits between NaN guard rows too: a read outside the map would poison an output
    _, fin = guarded((H, W), np.float32, np.nan)
    fin[:] = field
    field = fin
    out_full, out = guarded((H, W), np.float32, np.nan)
    surf_full, surf = guarded((H, W), np.float32, np.nan) if want_surf else (None, None)
    w = oracle.gauss_weights(R, sigma)
    cin = czero = czero_full = None
    if cm == 1:
        _, cin = guarded((H, W), np.uint32, 1)
        cin[:] = counts
        czero_full, czero = guarded((H, W), np.uint32, 7)   # the pass must retire every cell of the other buffer, and only those
    elif cm == 2:
        _, cin = guarded((H, W), np.uint8, 1)
        cin[:] = counts > 0
        czero_full, czero = guarded((H, W), np.uint8, 7)
    fn = {"rows": hostcheck.hc_gauss_rows}.get(kernel, hostcheck.hc_gauss_stream)
    rc = fn(P(field, C.c_float),
                                   None if cin is None else cin.ctypes.data_as(C.c_void_p),
                                   None if czero is None else czero.ctypes.data_as(C.c_void_p),
                                   P(out, C.c_float), None if surf is None else P(surf, C.c_float),
                                   C.c_int(W), C.c_int(H), C.c_int(chunk), C.c_int(R), P(w, C.c_float), C.c_int(cm), C.byref(p), C.c_int(1))
    assert rc == 0
    assert guards_intact(out_full, np.nan), "the kernel stored outside the output field"
    if surf_full is not None:
        assert guards_intact(surf_full, np.nan), "the kernel stored outside the sampler copy"
    if czero_full is not None:
        assert guards_intact(czero_full, 7), "the kernel retired deposit marks outside the map"
    return out, surf, czero


@pytest.fixture(params=[0, 1], ids=["scalar", "packed"])
def stream_packed(request, hostcheck):
    """Scalar taps, and the FFMA2 form (column taps on column pairs + the row taps whose operands are aligned pairs)."""
    hostcheck.hc_gauss_stream_set_packed(C.c_int(request.param))
    yield request.param
    hostcheck.hc_gauss_stream_set_packed(C.c_int(0))


def params_for(oracle, W, H, R, sigma, dep, rate=0.7):
    s = sm.init_preset_manager().get_preset("Default").settings.clone(
        blur_radius=float(R), blur_sigma=sigma, pheromone_diffusion_rate=rate, pheromone_deposition_amount=dep)
    return to_oracle_params(oracle, sm.SimSizeUniform.new(W, H, s.pheromone_decay_factor, s))


@pytest.mark.parametrize("R,sigma,W,H,chunk", [(1, 0.7, 288, 64, 64), (2, 1.0, 416, 200, 48), (3, 1.3, 516, 131, 40), (4, 2.0, 512, 96, 96),
                                               (5, 2.5, 1000, 97, 32), (6, 3.0, 384, 130, 56), (7, 3.5, 772, 65, 24), (8, 4.0, 640, 333, 104)])
def test_stream_diffuse_only_bits(oracle, hostcheck, stream_packed, R, sigma, W, H, chunk):
    p = params_for(oracle, W, H, R, sigma, dep=1.0)
    field = np.random.default_rng(R).random((H, W), dtype=np.float32)
    ref = oracle.trail_pass(field, p, counts=None, gauss_radius=R, gauss_sigma=sigma)
    got, _, _ = run_stream(hostcheck, oracle, field, p, R, sigma, chunk)
    assert bits_equal(got, ref), mismatch_report(got, ref, f"stream R={R}")


@pytest.mark.parametrize("cm,dep", [(1, 0.4), (1, 2.5), (2, 1.0)])
@pytest.mark.parametrize("R,sigma,W,H,chunk", [(1, 0.5, 300, 70, 32), (2, 1.0, 292, 64, 64), (4, 2.0, 520, 90, 48), (5, 2.5, 288, 100, 72), (8, 4.0, 548, 77, 40)])
def test_stream_full_step_bits(oracle, hostcheck, stream_packed, cm, dep, R, sigma, W, H, chunk):
    """Deposits merged by the pass (u32 counts, or u8 flags when dep >= 1), the other deposit buffer retired,
    the sampler copy written."""
    p = params_for(oracle, W, H, R, sigma, dep=dep)
    rng = np.random.default_rng(100 * R + cm)
    field = random_trail(W, H, seed=R, density=0.5)
    counts = (rng.random((H, W)) < 0.2).astype(np.uint32) * rng.integers(1, 4, (H, W)).astype(np.uint32)
    # deposits on the seams and in the corners
    counts[0, :5] = 1; counts[-1, -5:] = 2; counts[:3, -1] = 1; counts[-3:, 0] = 3
    ref = oracle.trail_pass(field, p, counts=counts.copy(), gauss_radius=R, gauss_sigma=sigma)
    got, surf, czero = run_stream(hostcheck, oracle, field, p, R, sigma, chunk, cm=cm, counts=counts, want_surf=True)
    assert bits_equal(got, ref), mismatch_report(got, ref, f"stream full step R={R} cm={cm}")
    assert bits_equal(surf, ref), "sampler copy differs from the row-major output"
    assert not czero.any(), "deposit marks of the next step's buffer were not all retired"


def test_stream_constant_field_and_mass(oracle, hostcheck):
    """Size-independent properties: a constant field stays constant under the blur (weights sum to 1 up to rounding),
    and with decay 0 and rate 1 the mass is conserved up to rounding."""
    W, H, R, sigma = 512, 128, 8, 4.0
    s = sm.init_preset_manager().get_preset("Default").settings.clone(blur_radius=float(R), blur_sigma=sigma, pheromone_diffusion_rate=1.0)
    p = to_oracle_params(oracle, sm.SimSizeUniform.new(W, H, 0.0, s))
    field = np.random.default_rng(0).random((H, W), dtype=np.float32)
    got, _, _ = run_stream(hostcheck, oracle, field, p, R, sigma, 64)
    assert abs(got.sum(dtype=np.float64) - field.sum(dtype=np.float64)) < 1e-5 * field.size
    const = np.full((H, W), 0.625, np.float32)
    got, _, _ = run_stream(hostcheck, oracle, const, p, R, sigma, 64)
    assert np.allclose(got, 0.625, rtol=0, atol=2e-7)


@pytest.mark.parametrize("R,sigma,strips", [(2, 1.0, 2), (5, 2.5, 3), (8, 4.0, 4)])
def test_stream_on_strips_bits(oracle, hostcheck, R, sigma, strips):
    """Strip mode of the kernel (multi-GPU diffusion-only passes, BASELINE config 5 at 2/4/8 GPUs): every strip runs the
    pass on its own rows with `ghost` rows of its neighbours above and below (no row wrap inside the buffer; the
    toroidal seam is the ring exchange), the ghosts are refreshed between passes -- the protocol of
    sm_diffuse_only on strips -- and the stitched result equals the single-domain oracle bit for bit."""
    W, H, ghost, passes = 320, 64 * strips, 9, 3
    p = params_for(oracle, W, H, R, sigma, dep=1.0)
    field = np.random.default_rng(strips).random((H, W), dtype=np.float32)
    ref = field
    for _ in range(passes):
        ref = oracle.trail_pass(ref, p, counts=None, gauss_radius=R, gauss_sigma=sigma)
    w = oracle.gauss_weights(R, sigma)
    rows = H // strips
    cur = field
    for _ in range(passes):
        new = np.empty_like(cur)
        for r in range(strips):
            y0 = r * rows
            buf = np.full((rows + 2 * ghost, W), np.nan, np.float32)      # [ghost | owned | ghost], filled by the "exchange"
            idx = np.arange(y0 - ghost, y0 + rows + ghost) % H
            buf[:] = cur[idx]
            # rows past the R the pass may touch stay poisoned: reading them would show up in the output
            buf[:ghost - R] = np.nan
            buf[ghost + rows + R:] = np.nan
            out = np.full((rows + 2 * ghost, W), np.nan, np.float32)
            pp = params_for(oracle, W, rows, R, sigma, dep=1.0)
            rc = hostcheck.hc_gauss_stream(P(buf[ghost:], C.c_float), None, None, P(out[ghost:], C.c_float), None,
                                           C.c_int(W), C.c_int(rows), C.c_int(40), C.c_int(R), P(w, C.c_float), C.c_int(0),
                                           C.byref(pp), C.c_int(0))
            assert rc == 0
            assert np.all(np.isnan(out[:ghost])) and np.all(np.isnan(out[ghost + rows:])), "stored into the ghost rows"
            new[y0:y0 + rows] = out[ghost:ghost + rows]
        cur = new
    assert bits_equal(cur, ref), mismatch_report(cur, ref, f"strips R={R}")


# ---- the register-streaming kernel (slime_mold_b200/csrc/gauss_rows.cuh; one halo lane per side up to radius 4, two above), same emulation ----
# widths: one warp exactly (128), a ragged last warp, several CTAs (> 480 columns), a last output lane next to the seam;
# chunk heights that are / are not multiples of the batch of 2R+1 rows, a ragged last chunk, one chunk for the whole map
@pytest.fixture(params=[0, 1, 2], ids=["scalar", "packed", "packed2"])
def rows_packed(request, hostcheck):
    """The forms of the taps: scalar FFMA; FFMA2 on column pairs for the column taps; that plus FFMA2 for the row taps whose
    operands are aligned pairs (on the host a packed lane IS the scalar fma: this checks the statement order)."""
    hostcheck.hc_gauss_rows_set_packed(C.c_int(request.param))
    yield request.param
    hostcheck.hc_gauss_rows_set_packed(C.c_int(0))


@pytest.mark.parametrize("R,sigma,W,H,chunk", [(1, 0.7, 128, 16, 16), (1, 0.5, 484, 50, 17), (2, 1.0, 416, 64, 48), (2, 1.3, 1000, 37, 16),
                                               (3, 1.3, 516, 61, 40), (3, 2.0, 240, 33, 33), (4, 2.0, 512, 48, 25), (4, 2.5, 964, 40, 16),
                                               (5, 2.5, 452, 40, 23), (6, 3.0, 448, 61, 61), (7, 3.5, 128, 30, 17), (8, 4.0, 900, 50, 34)])
def test_rows_diffuse_only_bits(oracle, hostcheck, rows_packed, R, sigma, W, H, chunk):
    p = params_for(oracle, W, H, R, sigma, dep=1.0)
    field = np.random.default_rng(R).random((H, W), dtype=np.float32)
    ref = oracle.trail_pass(field, p, counts=None, gauss_radius=R, gauss_sigma=sigma)
    got, _, _ = run_stream(hostcheck, oracle, field, p, R, sigma, chunk, kernel="rows")
    assert bits_equal(got, ref), mismatch_report(got, ref, f"rows R={R}")


@pytest.mark.parametrize("cm,dep", [(1, 0.4), (1, 2.5), (2, 1.0)])
@pytest.mark.parametrize("R,sigma,W,H,chunk", [(1, 0.5, 300, 40, 32), (2, 1.0, 292, 64, 20), (3, 1.5, 488, 30, 16), (4, 2.0, 520, 45, 45),
                                               (6, 3.0, 460, 36, 20), (8, 4.0, 448, 40, 40)])
def test_rows_full_step_bits(oracle, hostcheck, rows_packed, cm, dep, R, sigma, W, H, chunk):
    p = params_for(oracle, W, H, R, sigma, dep=dep)
    rng = np.random.default_rng(100 * R + cm)
    field = random_trail(W, H, seed=R, density=0.5)
    counts = (rng.random((H, W)) < 0.2).astype(np.uint32) * rng.integers(1, 4, (H, W)).astype(np.uint32)
    counts[0, :5] = 1; counts[-1, -5:] = 2; counts[:3, -1] = 1; counts[-3:, 0] = 3
    ref = oracle.trail_pass(field, p, counts=counts.copy(), gauss_radius=R, gauss_sigma=sigma)
    got, surf, czero = run_stream(hostcheck, oracle, field, p, R, sigma, chunk, cm=cm, counts=counts, want_surf=True, kernel="rows")
    assert bits_equal(got, ref), mismatch_report(got, ref, f"rows full step R={R} cm={cm}")
    assert bits_equal(surf, ref), "sampler copy differs from the row-major output"
    assert not czero.any(), "deposit marks of the next step's buffer were not all retired"


@pytest.mark.parametrize("R,sigma,strips", [(1, 0.8, 2), (2, 1.0, 3), (4, 2.0, 2), (7, 3.0, 2)])
def test_rows_on_strips_bits(oracle, hostcheck, R, sigma, strips):
    """Strip mode (no row wrap, ghost rows filled by the neighbours), as test_stream_on_strips_bits."""
    W, H, ghost, passes = 256, 32 * strips, 9, 3
    p = params_for(oracle, W, H, R, sigma, dep=1.0)
    field = np.random.default_rng(strips).random((H, W), dtype=np.float32)
    ref = field
    for _ in range(passes):
        ref = oracle.trail_pass(ref, p, counts=None, gauss_radius=R, gauss_sigma=sigma)
    w = oracle.gauss_weights(R, sigma)
    rows = H // strips
    cur = field
    for _ in range(passes):
        new = np.empty_like(cur)
        for r in range(strips):
            y0 = r * rows
            buf = cur[np.arange(y0 - ghost, y0 + rows + ghost) % H].copy()
            buf[:ghost - R] = np.nan                                       # rows the pass must not touch stay poisoned
            buf[ghost + rows + R:] = np.nan
            out = np.full((rows + 2 * ghost, W), np.nan, np.float32)
            pp = params_for(oracle, W, rows, R, sigma, dep=1.0)
            rc = hostcheck.hc_gauss_rows(P(buf[ghost:], C.c_float), None, None, P(out[ghost:], C.c_float), None,
                                         C.c_int(W), C.c_int(rows), C.c_int(20), C.c_int(R), P(w, C.c_float), C.c_int(0),
                                         C.byref(pp), C.c_int(0))
            assert rc == 0
            assert np.all(np.isnan(out[:ghost])) and np.all(np.isnan(out[ghost + rows:])), "stored into the ghost rows"
            new[y0:y0 + rows] = out[ghost:ghost + rows]
        cur = new
    assert bits_equal(cur, ref), mismatch_report(cur, ref, f"rows on strips R={R}")
