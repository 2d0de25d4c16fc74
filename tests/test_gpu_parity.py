"""Parity tests proper: the CUDA path (through the C ABI / CudaBackend) against the CPU
oracle on identical seeded inputs -- bit-exact on agents (x, y, angle, speed) and on
the trail.  -m gpu."""
import os

import numpy as np
import pytest

import slime_mold_b200 as sm
from conftest import bits_equal, mismatch_report
from presets_util import PRESET_NAMES, edge_agents, preset_uniform, random_trail, to_oracle_params

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def settings_for(name):
    return sm.init_preset_manager().get_preset(name).settings


def run_pair(oracle, name, W, H, N, steps, seed=3, trail=None, check_every=None, **backend_kw):
    s = settings_for(name)
    u = preset_uniform(name, W, H)
    p = to_oracle_params(oracle, u)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, seed)
    sim = oracle.Sim(p, ag, trail=trail)
    be = sm.CudaBackend.new(W, H, s, agent_count=N, **backend_kw)
    be.write_agents(ag)
    if trail is not None:
        be.write_trail(trail)
    done = 0
    for chunk in (check_every or [steps]):
        sim.step(chunk)
        be.step(chunk)
        done += chunk
        a = be.read_agents()
        t = be.read_trail()
        assert bits_equal(a, sim.agents), f"{name} step {done}: " + mismatch_report(a, sim.agents, "agents")
        assert bits_equal(t, sim.trail), f"{name} step {done}: " + mismatch_report(t, sim.trail, "trail")
    be.close()


@pytest.mark.parametrize("layout", ["auto", "tiled"])
@pytest.mark.parametrize("name", PRESET_NAMES)
def test_preset_one_and_many_steps(oracle, engine_lib, monkeypatch, name, layout):
    # 320x256 map, 40k agents: fast kernel (W % 4 == 0), sort every 16 steps (default); u8 deposit flags row-major (what a map
    # of this size gets) and in 8x8 tiles (what maps from 2^23 cells up get)
    monkeypatch.setenv("SM_FLAG_LAYOUT", layout)
    run_pair(oracle, name, 320, 256, 40000, 60, check_every=[1, 1, 18, 40])


@pytest.mark.parametrize("name", ["Default", "Snake", "Curls"])
def test_large_sensor_on_bigger_map(oracle, engine_lib, name):
    # sd = 225 (Snake) needs a map where the sensors actually land inside
    run_pair(oracle, name, 1024, 768, 200000, 20, trail=random_trail(1024, 768, seed=5), check_every=[1, 19])


@pytest.mark.parametrize("shape", [(37, 23), (5, 3), (8, 1), (4, 4), (12, 7), (1, 1), (2, 2), (1000, 3)])
def test_ragged_and_tiny_maps(oracle, engine_lib, shape):
    W, H = shape
    run_pair(oracle, "Default", W, H, 500, 8, trail=random_trail(W, H, seed=1, density=0.6), check_every=[1, 7])


def test_single_agent_and_sort_modes(oracle, engine_lib):
    run_pair(oracle, "Waves", 64, 64, 1, 10)
    run_pair(oracle, "Waves", 256, 128, 30000, 20, flags=sm.SM_FLAG_NO_SORT)
    run_pair(oracle, "Waves", 256, 128, 30000, 20, sort_interval=1)
    run_pair(oracle, "Waves", 256, 128, 30000, 20, sort_interval=3)


def test_device_init_matches_oracle(oracle, engine_lib):
    W, H, N = 640, 480, 100000
    s = settings_for("Threads")
    be = sm.CudaBackend.new(W, H, s, agent_count=N)
    be.init_agents(seed=77)
    a = be.read_agents()
    ref = oracle.init_agents(N, W, H, s.agent_speed_min, s.agent_speed_max, 77)
    assert bits_equal(a, ref), mismatch_report(a, ref, "init")
    be.step(5)
    p = to_oracle_params(oracle, preset_uniform("Threads", W, H))
    sim = oracle.Sim(p, ref)
    sim.step(5)
    assert bits_equal(be.read_agents(), sim.agents) and bits_equal(be.read_trail(), sim.trail)
    # reassign_agent_speeds (main.rs:101-145), seeded, on the device, after the agents were cell-sorted
    be.reassign_agent_speeds(seed=5)
    exp = sim.agents.copy()
    oracle.reassign_speeds(exp, s.agent_speed_min, s.agent_speed_max, 5)
    assert bits_equal(be.read_agents(), exp)
    be.close()


def test_partial_upload_download_and_trail_rectangles(oracle, engine_lib):
    W, H, N = 128, 96, 5000
    be = sm.CudaBackend.new(W, H, agent_count=N)
    ag = oracle.init_agents(N, W, H, 30, 50, 1)
    be.write_agents(ag)
    be.step(3)                                   # forces a sort: storage order != index order
    base = be.read_agents()
    patch = oracle.init_agents(700, W, H, 30, 50, 9)
    be.write_agents(patch, first=1200)
    got = be.read_agents()
    exp = base.copy(); exp[1200:1900] = patch
    assert bits_equal(got, exp)
    assert bits_equal(be.read_agents(first=1000, n=500), exp[1000:1500])
    # trail rectangles
    full = be.read_trail()
    rect = np.random.default_rng(0).random((10, 20), dtype=np.float32)
    be.write_trail(rect, x0=30, y0=40)
    exp_t = full.copy(); exp_t[40:50, 30:50] = rect
    assert bits_equal(be.read_trail(), exp_t)
    assert bits_equal(be.read_trail(x0=25, y0=38, w=40, h=20), exp_t[38:58, 25:65])
    be.clear_trail()
    assert not be.read_trail().any()
    with pytest.raises(sm.SlimeError):
        be.read_trail(x0=100, y0=0, w=100, h=1)
    with pytest.raises(sm.SlimeError):
        be.write_agents(patch, first=N - 10)
    be.close()


def test_agent_count_change_and_resize(oracle, engine_lib):
    W, H = 200, 120
    s = settings_for("Default")
    be = sm.CudaBackend.new(W, H, s, agent_count=1000)
    be.init_agents(1)
    be.step(2)
    be.set_agent_count(7000, seed=4)              # N key: everything re-randomised (main.rs:697-713)
    assert be.agent_count == 7000
    ref = oracle.init_agents(7000, W, H, s.agent_speed_min, s.agent_speed_max, 4)
    assert bits_equal(be.read_agents(), ref)
    be.step(4)
    before = be.read_agents()
    be.resize(320, 240)                           # main.rs:954-1015
    exp = before.copy(); oracle.rescale_agents(exp, W, H, 320, 240)
    assert bits_equal(be.read_agents(), exp)
    assert be.read_trail().shape == (240, 320) and not be.read_trail().any()
    p = to_oracle_params(oracle, preset_uniform("Default", 320, 240))
    sim = oracle.Sim(p, exp)
    sim.step(6); be.step(6)
    assert bits_equal(be.read_agents(), sim.agents) and bits_equal(be.read_trail(), sim.trail)
    be.close()


def test_parameter_update_mid_run(oracle, engine_lib):
    # update_settings (main.rs:83-99) between frames; preset switch does not touch the agents (:794-833)
    W, H, N = 256, 256, 30000
    be = sm.CudaBackend.new(W, H, settings_for("Default"), agent_count=N)
    ag = oracle.init_agents(N, W, H, 30, 50, 2)
    be.write_agents(ag)
    sim = oracle.Sim(to_oracle_params(oracle, preset_uniform("Default", W, H)), ag)
    sim.step(5); be.step(5)
    for name in ("Curls", "Mesh", "Sponge"):
        be.update_settings(settings_for(name))
        sim.p = to_oracle_params(oracle, preset_uniform(name, W, H))
        sim.step(5); be.step(5)
        assert bits_equal(be.read_agents(), sim.agents), name
        assert bits_equal(be.read_trail(), sim.trail), name
    be.close()


@pytest.mark.parametrize("dep", [0.05, 0.3, 2.5])
def test_fractional_and_large_deposit(oracle, engine_lib, dep):
    W, H, N = 128, 128, 60000            # ~3.7 agents per cell: multi-deposit cells are common
    s = settings_for("Default").clone(pheromone_deposition_amount=dep, pheromone_decay_factor=30.0)
    u = sm.SimSizeUniform.new(W, H, s.pheromone_decay_factor, s)
    ag = oracle.init_agents(N, W, H, 30, 50, 8)
    sim = oracle.Sim(to_oracle_params(oracle, u), ag)
    be = sm.CudaBackend.new(W, H, s, agent_count=N)
    be.write_agents(ag)
    sim.step(12); be.step(12)
    assert bits_equal(be.read_agents(), sim.agents) and bits_equal(be.read_trail(), sim.trail)
    be.close()


@pytest.mark.parametrize("shape", [(4096, 64), (1920, 1080), (333, 77), (8, 8)])
def test_diffuse_only(oracle, engine_lib, shape):
    W, H = shape
    u = preset_uniform("Default", W, H)
    p = to_oracle_params(oracle, u)
    field = np.random.default_rng(W).random((H, W), dtype=np.float32)
    be = sm.CudaBackend.new(W, H, agent_count=1)
    be.write_trail(field)
    be.diffuse_only(3)
    ref = field
    for _ in range(3):
        ref = oracle.trail_pass(ref, p, counts=None)
    got = be.read_trail()
    assert bits_equal(got, ref), mismatch_report(got, ref, "diffuse_only")
    st = be.trail_statistics()
    assert abs(st.sum - ref.sum(dtype=np.float64)) < 1e-6 * ref.size
    assert st.max == ref.max() and st.nonzero == np.count_nonzero(ref)
    be.close()


@pytest.mark.parametrize("R,sigma", [(1, 0.5), (2, 1.0), (4, 2.0), (8, 4.0)])
def test_gaussian_extension(oracle, engine_lib, R, sigma):
    # EXTENSION, no reference semantics: compared with the oracle's definition only
    W, H = 200, 150
    s = settings_for("Default").clone(blur_radius=float(R), blur_sigma=sigma, pheromone_diffusion_rate=0.6)
    u = sm.SimSizeUniform.new(W, H, s.pheromone_decay_factor, s)
    p = to_oracle_params(oracle, u)
    field = np.random.default_rng(R).random((H, W), dtype=np.float32)
    be = sm.CudaBackend.new(W, H, s, agent_count=1, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
    be.write_trail(field)
    be.diffuse_only(2)
    ref = field
    for _ in range(2):
        ref = oracle.trail_pass(ref, p, counts=None, gauss_radius=R, gauss_sigma=sigma)
    got = be.read_trail()
    assert bits_equal(got, ref), mismatch_report(got, ref, "gauss")
    be.close()


@pytest.mark.parametrize("kernel", ["scalar", "two_pass", "stream", "rows", "rows_packed2"])
@pytest.mark.parametrize("R,sigma,W,H", [(1, 0.7, 160, 64), (2, 1.0, 416, 200), (3, 1.3, 517, 131), (4, 2.0, 256, 96), (5, 2.5, 1000, 97),
                                         (6, 3.0, 384, 130), (7, 3.5, 772, 65), (8, 4.0, 640, 333)])
def test_gaussian_extension_fused_and_two_pass(oracle, engine_lib, monkeypatch, R, sigma, W, H, kernel):
    """EXTENSION: the fused shared-memory kernels (FFMA2-packed and scalar; ragged tiles, wrap on all four sides), the
    streaming kernel (maps of at least 288 x 64 with W % 4 == 0; smaller ones fall back to the tile kernel) and the
    two-pass form give the oracle's bits, in diffusion-only passes and in full steps (deposit counts merged by the pass)."""
    monkeypatch.setenv("SM_GAUSS_TWO_PASS", "1" if kernel == "two_pass" else "0")
    monkeypatch.setenv("SM_GAUSS_KERNEL", "rows" if kernel.startswith("rows") else "stream" if kernel.startswith("stream") else "tile")
    monkeypatch.setenv("SM_GAUSS_ROWS_PACKED", "2" if kernel == "rows_packed2" else "0")    # FFMA2 form of the taps
    s = settings_for("Default").clone(blur_radius=float(R), blur_sigma=sigma, pheromone_diffusion_rate=0.7, pheromone_deposition_amount=0.4)
    u = sm.SimSizeUniform.new(W, H, s.pheromone_decay_factor, s)
    p = to_oracle_params(oracle, u)
    N = 20_000
    field = random_trail(W, H, seed=R)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 4)
    be = sm.CudaBackend.new(W, H, s, agent_count=N, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
    be.write_trail(field)
    be.write_agents(ag)
    be.diffuse_only(2)
    ref = field
    for _ in range(2):
        ref = oracle.trail_pass(ref, p, counts=None, gauss_radius=R, gauss_sigma=sigma)
    got = be.read_trail()
    assert bits_equal(got, ref), mismatch_report(got, ref, "gauss diffuse-only")
    # full steps: agents (phase split) -> counts -> Gaussian pass
    be.step(3)
    a = ag.copy()
    counts = np.zeros((H, W), np.uint32)
    for _ in range(3):
        oracle.agents_phase_split(a, ref, counts, p)
        ref = oracle.trail_pass(ref, p, counts=counts, gauss_radius=R, gauss_sigma=sigma)
    assert bits_equal(be.read_agents(), a), "agents"
    got = be.read_trail()
    assert bits_equal(got, ref), mismatch_report(got, ref, "gauss full step")
    be.close()


@pytest.mark.parametrize("kernel", ["stream", "rows", "rows_packed2"])
@pytest.mark.parametrize("chunk", [0, 40])
@pytest.mark.parametrize("dep", [1.0, 0.4])
@pytest.mark.parametrize("R,sigma,W,H", [(1, 0.7, 288, 64), (2, 1.0, 512, 256), (3, 1.3, 516, 131), (4, 2.0, 1024, 96), (6, 3.0, 388, 150),
                                         (8, 4.0, 1280, 333)])
def test_gaussian_stream_full_steps(oracle, engine_lib, monkeypatch, R, sigma, W, H, dep, chunk, kernel):
    """EXTENSION, streaming kernel (gauss_stream.cuh): full steps in Gaussian mode -- with dep >= 1 the agents mark u8
    deposit flags and sense through the texture copy the pass keeps in step (the box-blur step's fast path), with a
    fractional deposit they count -- against the oracle's phase-split agents + Gaussian pass; CTA chunk heights
    chosen by the engine and forced to a value that leaves ragged chunks."""
    monkeypatch.setenv("SM_GAUSS_KERNEL", "rows" if kernel.startswith("rows") else "stream")
    monkeypatch.setenv("SM_GAUSS_ROWS_PACKED", "2" if kernel == "rows_packed2" else "0")
    monkeypatch.setenv("SM_GAUSS_CHUNK", str(chunk))
    s = settings_for("Default").clone(blur_radius=float(R), blur_sigma=sigma, pheromone_diffusion_rate=0.8, pheromone_deposition_amount=dep)
    u = sm.SimSizeUniform.new(W, H, s.pheromone_decay_factor, s)
    p = to_oracle_params(oracle, u)
    N = 60_000
    field = random_trail(W, H, seed=10 + R)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 6)
    be = sm.CudaBackend.new(W, H, s, agent_count=N, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
    be.write_trail(field)
    be.write_agents(ag)
    a, ref = ag.copy(), field
    counts = np.zeros((H, W), np.uint32)
    for chunk_steps in (1, 4):
        be.step(chunk_steps)
        for _ in range(chunk_steps):
            oracle.agents_phase_split(a, ref, counts, p)
            ref = oracle.trail_pass(ref, p, counts=counts, gauss_radius=R, gauss_sigma=sigma)
        assert bits_equal(be.read_agents(), a), "agents"
        got = be.read_trail()
        assert bits_equal(got, ref), mismatch_report(got, ref, "gauss stream full step")
    # a diffusion-only pass in between (marks the sampler copy stale), then steps again
    be.diffuse_only(1)
    ref = oracle.trail_pass(ref, p, counts=None, gauss_radius=R, gauss_sigma=sigma)
    be.step(2)
    for _ in range(2):
        oracle.agents_phase_split(a, ref, counts, p)
        ref = oracle.trail_pass(ref, p, counts=counts, gauss_radius=R, gauss_sigma=sigma)
    assert bits_equal(be.read_agents(), a), "agents after the mixed sequence"
    got = be.read_trail()
    assert bits_equal(got, ref), mismatch_report(got, ref, "gauss stream mixed sequence")
    st = be.trail_statistics()
    assert abs(st.sum - ref.sum(dtype=np.float64)) < 1e-6 * ref.size and st.max == ref.max()
    be.close()


@pytest.mark.parametrize("name", PRESET_NAMES)
def test_golden_vectors(engine_lib, name):
    g = np.load(os.path.join(GOLD, "golden_" + name.lower().replace(" ", "_") + ".npz"))
    H, W = g["trail1"].shape
    be = sm.CudaBackend.new(W, H, settings_for(name), agent_count=g["agents0"].shape[0])
    assert bytes(be.read_uniform()) == g["params"].tobytes()
    be.write_agents(g["agents0"])
    be.step(1)
    assert bits_equal(be.read_agents(), g["agents1"]) and bits_equal(be.read_trail(), g["trail1"])
    be.step(24)
    assert bits_equal(be.read_agents(), g["agents25"]) and bits_equal(be.read_trail(), g["trail25"])
    be.clear_trail(); be.write_trail(g["field"]); be.diffuse_only(1)
    assert bits_equal(be.read_trail(), g["diffused"])
    be.close()


def test_error_behaviour(engine_lib):
    with pytest.raises(sm.SlimeError):
        sm.CudaBackend.new(0, 10, agent_count=1)
    be = sm.CudaBackend.new(32, 32, agent_count=10)
    with pytest.raises(sm.SlimeError):
        be.step(1)                                # agents never initialised
    bad = sm.SimSizeUniform.new(64, 32, 10.0, sm.Settings.default())
    with pytest.raises(sm.SlimeError):
        be.write_uniform(bad)                     # size mismatch must be an error, not a silent resize
    with pytest.raises(sm.SlimeError):
        be.comm_init(b"\0" * 128)
    be.close()


@pytest.mark.parametrize("name", ["Default", "Waves", "Mesh"])
def test_edge_agents(oracle, engine_lib, name):
    """Rare paths on the device: +-0 headings with jitter 0 (dead-hash shortcut), huge headings (f64
    reduction, library fmod), positions outside the map / on the seams, non-finite state."""
    W, H = 160, 96
    ag = edge_agents(W, H)
    trail = random_trail(W, H, seed=6, density=0.7)
    u = preset_uniform(name, W, H)
    sim = oracle.Sim(to_oracle_params(oracle, u), ag, trail=trail)
    be = sm.CudaBackend.new(W, H, settings_for(name), agent_count=len(ag), flags=sm.SM_FLAG_NO_SORT)
    be.write_agents(ag)
    be.write_trail(trail)
    for step in range(3):
        sim.step(1); be.step(1)
        a = be.read_agents()
        assert bits_equal(a, sim.agents), (step, mismatch_report(a, sim.agents, "agents"))
        assert bits_equal(be.read_trail(), sim.trail), step
    be.close()


def test_deposit_mode_switching_and_negative_trail(oracle, engine_lib):
    """The engine marks deposits with u8 flags when dep >= 1 and the field is non-negative, and counts
    them otherwise; switching back and forth mid-run (dep 1.0 -> 0.3 -> 2.0 -> 0.05) and starting from a
    field with negative / NaN cells must not change a bit."""
    W, H, N = 192, 160, 50000
    s = settings_for("Default")
    ag = oracle.init_agents(N, W, H, s.agent_speed_min, s.agent_speed_max, 21)
    trail = random_trail(W, H, seed=9, density=0.8)
    trail[5:20, 7:90] = -0.25          # negative cells: clamp(t + dep, 0, 1) != 1 for dep = 1 only if t < 0 ...
    trail[40, 40] = np.nan
    trail[41, 41] = -3.0               # ... e.g. -3 + 1 -> clamp -> 0
    be = sm.CudaBackend.new(W, H, s, agent_count=N)
    be.write_agents(ag)
    be.write_trail(trail)
    sim = oracle.Sim(to_oracle_params(oracle, preset_uniform("Default", W, H)), ag, trail=trail)
    for dep in (1.0, 1.0, 0.3, 0.3, 2.0, 1.0, 0.05, 1.0, 1.0):
        ss = s.clone(pheromone_deposition_amount=dep)
        be.update_settings(ss)
        sim.p = to_oracle_params(oracle, sm.SimSizeUniform.new(W, H, ss.pheromone_decay_factor, ss))
        sim.step(3); be.step(3)
        a, t = be.read_agents(), be.read_trail()
        assert bits_equal(a, sim.agents), (dep, mismatch_report(a, sim.agents, "agents"))
        assert bits_equal(t, sim.trail), (dep, mismatch_report(t, sim.trail, "trail"))
    be.close()


@pytest.mark.parametrize("dep", [1.0, 0.3])
def test_fused_step_statistics(oracle, engine_lib, dep):
    """sm_trail_statistics after sm_step comes from the trail pass itself (fused reduction; flags and counts
    deposit modes), after anything else from a sweep of its own: both must describe the field read back."""
    W, H, N = 512, 384, 150_000
    s = settings_for("Default").clone(pheromone_deposition_amount=dep)
    be = sm.CudaBackend.new(W, H, s, agent_count=N, device=0)
    be.init_agents(seed=3)

    def check():
        st = be.trail_statistics()
        t = be.read_trail()
        t64 = t.astype(np.float64)
        # fused path: row quads are combined in f32 before they enter the f64 accumulators (relative error < 2e-7 per quad)
        assert abs(st.sum - t64.sum()) <= 2e-7 * max(t64.sum(), 1.0), (st.sum, t64.sum())
        assert abs(st.sum_sq - (t64 * t64).sum()) <= 2e-7 * max((t64 * t64).sum(), 1.0)
        assert st.max == t.max() and st.nonzero == np.count_nonzero(t)
        return st

    for n in (1, 7, 16, 1):                      # crosses a sort
        be.step(n)
        a = check()
        b = check()                              # reading twice does not disturb the accumulator
        assert (a.max, a.nonzero) == (b.max, b.nonzero)
        assert abs(a.sum - b.sum) <= 1e-12 * max(a.sum, 1.0)      # (the separate sweep sums its blocks in atomic order)
    be.diffuse_only(2)
    check()                                      # not a full-step pass: separate sweep
    be.write_trail(random_trail(W, H, seed=8))
    check()
    be.step(3)
    check()
    be.clear_trail()
    st = check()
    assert st.sum == 0.0 and st.nonzero == 0
    be.close()


# ---- parameter changes between steps (sm_set_params: the reference's hold-key + arrow paths, main.rs:407-433) -------------
def test_sensor_distance_changes_between_steps(oracle, engine_lib):
    W, H, N = 640, 512, 80000
    s0 = settings_for("Default")
    ag = oracle.init_agents(N, W, H, s0.agent_speed_min, s0.agent_speed_max, 21)
    tr = random_trail(W, H, seed=2)
    sim = oracle.Sim(to_oracle_params(oracle, preset_uniform("Default", W, H)), ag, trail=tr)
    be = sm.CudaBackend.new(W, H, s0, agent_count=N)
    be.write_agents(ag); be.write_trail(tr)
    for sd in (20.0, 225.0, 47.0, 48.0, 100.0, 20.0):
        s = s0.clone(agent_sensor_distance=sd, agent_sensor_angle=1.1)
        be.update_settings(s)
        sim.p = to_oracle_params(oracle, sm.SimSizeUniform.new(W, H, s.pheromone_decay_factor, s))
        sim.step(3); be.step(3)
        a, t = be.read_agents(), be.read_trail()
        assert bits_equal(a, sim.agents), f"sd {sd}: " + mismatch_report(a, sim.agents, "agents")
        assert bits_equal(t, sim.trail), f"sd {sd}: " + mismatch_report(t, sim.trail, "trail")
    be.close()


@pytest.mark.parametrize("pairs", ["0", "1"])
@pytest.mark.parametrize("W,H", [(520, 300), (8, 1), (256, 7), (136, 33)])
def test_sampler_copy_written_row_by_row_or_in_pairs(oracle, engine_lib, monkeypatch, pairs, W, H):
    # the trail pass writes the sampler's block-linear copy as whole sectors (row pairs, SM_SURF_PAIRS=1, default) or row by row;
    # odd row counts, a single row, a ragged last warp
    monkeypatch.setenv("SM_SURF_PAIRS", pairs)
    run_pair(oracle, "Waves", W, H, 3000, 9, trail=random_trail(W, H, seed=8), check_every=[1, 8])


# ---- u8 deposit flags: 8 x 8-cell tiles (default on one GPU from 2^23 cells up, forced here on small maps) or row-major --------------------
@pytest.mark.parametrize("layout", ["tiled", "linear"])
@pytest.mark.parametrize("name,W,H", [("Default", 512, 264), ("Waves", 8, 8), ("Firecracker Trees", 1032, 16), ("Curls", 136, 40),
                                      ("Default", 260, 136), ("Default", 256, 100)])
def test_deposit_flag_layouts(oracle, engine_lib, monkeypatch, layout, name, W, H):
    """kernels.cuh flag_tile_offset: the agent kernel marks cells in 64-byte tiles, the trail pass and the display pass read
    them back; the last two shapes are not whole tiles and stay row-major.  Includes a deposit-mode change (flags -> counts
    -> flags) and a frame drawn from the step's flags."""
    monkeypatch.setenv("SM_FLAG_LAYOUT", layout)
    N = max(64, W * H // 3)
    s = settings_for(name)
    u = preset_uniform(name, W, H)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 21)
    tr = random_trail(W, H, seed=6)
    sim = oracle.Sim(to_oracle_params(oracle, u), ag, trail=tr)
    be = sm.CudaBackend.new(W, H, s, agent_count=N, sort_interval=5)
    be.write_agents(ag)
    be.write_trail(tr)
    for dep, n in ((None, 7), (0.4, 3), (None, 6)):
        ss = s if dep is None else s.clone(pheromone_deposition_amount=dep)
        be.update_settings(ss)
        uu = sm.SimSizeUniform.new(W, H, ss.pheromone_decay_factor, ss)
        sim.p = to_oracle_params(oracle, uu)
        sim.step(n - 1); be.step(n - 1)
        # the last step of the leg by hand, for the frame the reference draws between decay and diffuse
        oracle.agents_phase_split(sim.agents, sim.trail, sim.counts, sim.p)
        pre = sim.trail.copy()
        oracle.deposit_merge(pre, sim.counts, ss.pheromone_deposition_amount)
        oracle.decay(pre, uu.decay_factor)
        sim.trail = oracle.diffuse(pre, uu.diffusion_rate)
        be.step(1)
        lut = np.random.default_rng(2).integers(0, 256, 768).astype(np.uint8)
        be.set_lut(lut)
        frame = be.render(W + 9, H + 5)
        want = oracle.display(pre, lut, W + 9, H + 5)
        assert np.array_equal(frame, want), f"dep {dep}: {np.count_nonzero(frame != want)} frame bytes differ"
        a, t = be.read_agents(), be.read_trail()
        assert bits_equal(a, sim.agents), f"dep {dep}: " + mismatch_report(a, sim.agents, "agents")
        assert bits_equal(t, sim.trail), f"dep {dep}: " + mismatch_report(t, sim.trail, "trail")
    be.close()


# ---- CUDA-graph replay of whole sort periods (engine.cu: graph_steps) ------------------------------------------------------
@pytest.mark.parametrize("name,sort_interval", [("Default", 0), ("Waves", 5), ("Snake", 7)])
def test_step_graph_equals_launch_path_and_oracle(oracle, engine_lib, monkeypatch, name, sort_interval):
    """sm_step(n) replays captured periods (2 x sort_interval steps) when n is large enough; parameter changes, uploads and
    statistics requests in between fall back to the launch path and re-capture.  Same bits as the oracle throughout, and as
    an engine with the graphs switched off."""
    W, H, N = 384, 256, 70000
    if name == "Waves":
        monkeypatch.setenv("SM_FLAG_LAYOUT", "tiled")     # the captured kernels include the tiled-flag instantiations
    s = settings_for(name)
    u = preset_uniform(name, W, H)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 13)
    sim = oracle.Sim(to_oracle_params(oracle, u), ag)
    be = sm.CudaBackend.new(W, H, s, agent_count=N, sort_interval=sort_interval)
    be.write_agents(ag)
    monkeypatch.setenv("SM_STEP_GRAPH", "0")
    plain = sm.CudaBackend.new(W, H, s, agent_count=N, sort_interval=sort_interval)
    monkeypatch.delenv("SM_STEP_GRAPH")
    plain.write_agents(ag)
    si = sort_interval or 24
    plan = [3, 2 * si + 1, 4 * si, 1, 6 * si + 7]
    for k, n in enumerate(plan):
        if k == 2:                                   # a parameter change between two graph launches
            s = s.clone(agent_turn_speed=0.9, pheromone_decay_factor=25.0)
            u = sm.SimSizeUniform.new(W, H, s.pheromone_decay_factor, s)
            sim.p = to_oracle_params(oracle, u)
            be.update_settings(s); plain.update_settings(s)
        if k == 3:
            be.trail_statistics(); plain.trail_statistics()      # arms the fused statistics: launch path for a while
        if k == 4:                                   # an upload in between: the sampler copy is stale, the order is reset
            tr = random_trail(W, H, seed=9)
            sim.trail = tr.copy(); be.write_trail(tr); plain.write_trail(tr)
        sim.step(n); be.step(n); plain.step(n)
        a, t = be.read_agents(), be.read_trail()
        assert bits_equal(a, sim.agents), f"{name} leg {k}: " + mismatch_report(a, sim.agents, "agents")
        assert bits_equal(t, sim.trail), f"{name} leg {k}: " + mismatch_report(t, sim.trail, "trail")
        assert bits_equal(plain.read_agents(), a) and bits_equal(plain.read_trail(), t)
    tg, tp = be.timing(), plain.timing()
    assert tg.steps == tp.steps == sum(plan) and tg.kernel_launches == tp.kernel_launches       # replays are counted like launches
    be.close(); plain.close()
