"""The engine's __host__ __device__ arithmetic (device_math.cuh / agent_core.cuh /
trail_core.cuh), instantiated for the host by tests/hostcheck (TEST ONLY), against the
oracle -- bit for bit.  This is how kernel arithmetic is debugged without a GPU; the
device instantiation is checked by the -m gpu tests."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal
from presets_util import PRESET_NAMES, edge_agents, preset_uniform, random_trail, to_oracle_params


def P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def special_values():
    return np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 8192.0, -8192.0, 8192.001, 1e15, 2e15, 3e38, 1e-40,
                     6.2831855, -6.2831855, 12.566371], dtype=np.float32)


def test_sincos_bits(oracle, hostcheck):
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.integers(0, 2**32, 1_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32),
        rng.uniform(-10, 10, 500_000).astype(np.float32), rng.uniform(-9000, 9000, 300_000).astype(np.float32),
        rng.uniform(-2e10, 2e10, 300_000).astype(np.float32), special_values()])
    s0, c0 = oracle.sincos(x)
    s1, c1 = np.empty_like(x), np.empty_like(x)
    hostcheck.hc_sincos_array(P(x, C.c_float), P(s1, C.c_float), P(c1, C.c_float), C.c_uint64(x.size))
    assert bits_equal(s0, s1) and bits_equal(c0, c1)


@pytest.mark.parametrize("b", [6.2831855, 1920.0, 1080.0, 4096.0, 65536.0, 1.0, 7.0, 37.0])
def test_fmod_exact_bits(oracle, hostcheck, b):
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.uniform(-3 * b, 3 * b, 400_000), rng.uniform(-1e7, 1e7, 200_000), rng.uniform(-1e12, 1e12, 50_000)]).astype(np.float32)
    a = np.concatenate([a, rng.integers(0, 2**32, 400_000, dtype=np.uint64).astype(np.uint32).view(np.float32), special_values(),
                        np.float32(b) * np.arange(-5, 6, dtype=np.float32)])
    bb = np.full_like(a, np.float32(b))
    r1 = np.empty_like(a)
    hostcheck.hc_fmod_array(P(a, C.c_float), P(bb, C.c_float), P(r1, C.c_float), C.c_uint64(a.size))
    assert bits_equal(oracle.fmod(a, bb), r1)


def test_div9_bits(oracle, hostcheck):
    rng = np.random.default_rng(2)
    d = np.concatenate([rng.integers(0, 2**32, 1_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32),
                        rng.uniform(0, 9, 500_000).astype(np.float32), special_values()])
    # domain of the fast path: the 9-tap sum starts from +0.0 and adds values >= +0, so it is never -0.0
    d = d[~((d == 0) & np.signbit(d))]
    r1 = np.empty_like(d)
    hostcheck.hc_div9_array(P(d, C.c_float), P(r1, C.c_float), C.c_uint64(d.size))
    assert bits_equal(oracle.div9(d), r1)


def test_hash_and_init_bits(oracle, hostcheck):
    rng = np.random.default_rng(3)
    n = 500_000
    idx = rng.integers(0, 2**31 - 1, n).astype(np.int32)
    x = rng.uniform(0, 32768, n).astype(np.float32)
    y = rng.uniform(0, 32768, n).astype(np.float32)
    r1 = np.empty_like(x)
    hostcheck.hc_hash01_array(P(idx, C.c_int32), P(x, C.c_float), P(y, C.c_float), P(r1, C.c_float), C.c_uint64(n))
    assert bits_equal(oracle.hash01(idx, x, y), r1)
    a0 = oracle.init_agents(50_000, 1920, 1080, 30, 50, 7, first_id=123)
    a1 = np.empty_like(a0)
    hostcheck.hc_init_agents(P(a1, C.c_float), C.c_uint64(123), C.c_uint64(50_000), C.c_uint32(1920), C.c_uint32(1080),
                             C.c_float(30), C.c_float(50), C.c_uint64(7))
    assert bits_equal(a0, a1)


@pytest.mark.parametrize("name", PRESET_NAMES)
def test_step_loop_bits(oracle, hostcheck, name):
    W, H, N, steps = 320, 256, 20000, 12
    u = preset_uniform(name, W, H)
    p = to_oracle_params(oracle, u)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 3)
    sim = oracle.Sim(p, ag)
    a = ag.copy()
    tr = np.zeros((H, W), np.float32)
    cn = np.zeros((H, W), np.uint32)
    out = np.empty_like(tr)
    for step in range(steps):
        sim.step(1)
        hostcheck.hc_agents_phase_split(P(a, C.c_float), None, C.c_uint64(N), P(tr, C.c_float), P(cn, C.c_uint32), C.byref(p))
        hostcheck.hc_trail_pass(P(tr, C.c_float), P(cn, C.c_uint32), P(out, C.c_float), C.byref(p))
        tr, out = out, tr
        assert bits_equal(a, sim.agents), f"agents differ at step {step}"
        assert bits_equal(tr, sim.trail), f"trail differs at step {step}"


@pytest.mark.parametrize("name", ["Default", "Waves", "Mesh"])
def test_edge_agents_bits(oracle, hostcheck, name):
    """Rare paths: +-0 headings with jitter 0 (the dead-hash shortcut must not change a bit), huge
    headings (f64 reduction, library fmod), out-of-map and non-finite state."""
    W, H = 160, 96
    u = preset_uniform(name, W, H)
    p = to_oracle_params(oracle, u)
    ag = edge_agents(W, H)
    trail = random_trail(W, H, seed=6, density=0.7)
    a0 = ag.copy(); c0 = np.zeros((H, W), np.uint32)
    oracle.agents_phase_split(a0, trail, c0, p)
    a1 = ag.copy(); c1 = np.zeros((H, W), np.uint32)
    hostcheck.hc_agents_phase_split(P(a1, C.c_float), None, C.c_uint64(len(ag)), P(trail, C.c_float), P(c1, C.c_uint32), C.byref(p))
    assert bits_equal(a0, a1), [(i, ag[i], a0[i], a1[i]) for i in range(len(ag)) if not bits_equal(a0[i], a1[i])][:4]
    assert np.array_equal(c0, c1)


# ---- u8 deposit flags in 8 x 8-cell tiles (trail_core.cuh flag_tile_offset) -------------------------
def _tile_offsets(hostcheck, W, ys, wrap):
    xs = np.arange(W, dtype=np.int64)
    X, Y = np.meshgrid(xs, np.asarray(ys, dtype=np.int64))
    x, y = np.ascontiguousarray(X.ravel()), np.ascontiguousarray(Y.ravel())
    out = np.empty_like(x)
    hostcheck.hc_flag_tile_offsets(P(x, C.c_int64), P(y, C.c_int64), P(out, C.c_int64), C.c_uint64(x.size), C.c_int64(W), C.c_int64(wrap))
    return out.reshape(len(ys), W)


@pytest.mark.parametrize("W,H", [(8, 8), (64, 24), (520, 40), (1032, 16)])
def test_flag_tiles_one_gpu_layout(hostcheck, W, H):
    """One GPU (wrap = H): a permutation of the W*H bytes; the four rows a batch of the trail pass requests (y+1 .. y+4 with
    y % 4 == 0, across the toroidal seam too) are ONE aligned 32-byte sector per tile, row r of the batch at bytes 8r .. 8r+7 --
    what the lane-pair load of k_trail_rows assumes; the two prologue rows are the second half of the sector above; the
    region a chunk of 4 / 8 rows zeroes up front is the cells of rows y_begin+1 .. y_begin+n."""
    off = _tile_offsets(hostcheck, W, range(H), H)
    assert np.array_equal(np.sort(off.ravel()), np.arange(W * H))
    for y in range(0, H, 4):
        rows = [(y + 1 + u) % H for u in range(4)]
        for tx in range(W // 8):
            o = off[rows, tx * 8:(tx + 1) * 8]
            base = int(o.min())
            assert base % 32 == 0 and np.array_equal(o - base, np.arange(32).reshape(4, 8))
        # prologue of a chunk that starts at y (rows y-1 and y): second half of the sector that starts at y' = y - 4
        pro = off[[(y - 1) % H, y % H], :8]
        sector_start = int(off[(y - 3) % H, 0])              # row y-3 has y' = y - 4: first row of that sector
        assert sector_start % 32 == 0 and np.array_equal(pro - sector_start - 16, np.arange(16).reshape(2, 8))
    for n in (4, 8):
        for y0 in range(0, H, n):
            want = np.sort(off[[(y0 + 1 + u) % H for u in range(n)], :].ravel())
            # per tile: n*8 contiguous bytes from the tile's start + (y0 & 7) * 8
            tiles = (y0 >> 3) * W * 8 + np.arange(W // 8)[:, None] * 64 + (y0 & 7) * 8 + np.arange(n * 8)[None, :]
            assert np.array_equal(want, np.sort(tiles.ravel()))


@pytest.mark.parametrize("W,rows,G", [(64, 32, 48), (256, 64, 80), (16, 8, 8)])
def test_flag_tiles_strip_layout(hostcheck, W, rows, G):
    """Strips (wrap = 0): rows [-G, rows + G) relative to owned row 0 (G = ghost + pad, a multiple of 8) land inside the strip's
    buffer without collisions; owned row 0 sits at y' = -1, in the ghost tile row above; the rows barrier 1 pulls (-1 and `rows`)
    are W/8 eight-byte pieces 64 bytes apart (k_barrier_pull_tiled)."""
    ys = list(range(-G + 1, rows + G))                       # row -G itself would need y' = -G-1 (outside); nothing addresses it
    off = _tile_offsets(hostcheck, W, ys, 0)
    assert off.min() >= -G * W and off.max() < (rows + G) * W
    assert np.unique(off).size == off.size
    row = {y: off[i] for i, y in enumerate(ys)}
    assert row[0].max() < 0 and row[1].min() >= 0            # owned row 0 in the ghost tile row, row 1 opens the owned tiles
    for y in (-1, rows):
        r = row[y]
        assert np.array_equal(r.reshape(W // 8, 8) - r[0], np.arange(W // 8)[:, None] * 64 + np.arange(8)[None, :])


def test_flag_tiles_lane_pair_exchange():
    """The SHFL exchange of k_trail_rows' tiled path, restated: lanes 2j / 2j+1 own columns 0-3 / 4-7 of a tile, fetch the first /
    second 16 bytes of the batch's sector and must end up with their own four columns of each of the four rows."""
    rng = np.random.default_rng(4)
    sector = rng.integers(0, 2, 32).astype(np.uint8)          # 4 rows x 8 columns
    rows = sector.reshape(4, 8)
    q = {0: sector[:16].view(np.uint32), 1: sector[16:].view(np.uint32)}       # q.x .. q.w of the even / odd lane
    send = {lane: ((q[lane][0], q[lane][2]) if lane else (q[lane][1], q[lane][3])) for lane in (0, 1)}
    got = {}
    for lane in (0, 1):
        ra, rb = send[1 - lane]
        got[lane] = [ra, rb, q[lane][1], q[lane][3]] if lane else [q[lane][0], q[lane][2], ra, rb]
    for lane in (0, 1):
        for r in range(4):
            assert got[lane][r] == rows[r, 4 * lane:4 * lane + 4].view(np.uint32)[0]
