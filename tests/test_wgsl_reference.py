"""The oracle pinned to the REFERENCE'S OWN SHADER SOURCE.

tests/golden/wgsl_*.npz hold outputs of /root/reference/src/compute.wgsl and display.wgsl, executed by the WGSL
interpreter of tests/wgsl_interp.py (tests/golden/make_wgsl_golden.py; each file carries the SHA-256 of the shader text).
Here:
  * the C oracle must reproduce them BIT FOR BIT -- `so_step_sequential` against the sequential schedule (agents in
    index order, in-place raster diffuse), `so_step_phase_split` against the lockstep schedule (dep >= 1), the display
    restatement against display.wgsl -- edge-case agents included;
  * against the same shader run with numpy's libm sin / cos (a different conforming backend) the oracle stays inside the
    stated tolerance;
  * when /root/reference is present (build container) the shader is re-run and must reproduce the committed vectors,
    and the digest of the shader text must match: the fixtures really are outputs of that file.
The GPU-side counterpart is tests/test_gpu_wgsl.py.
"""
import os
import sys

import numpy as np
import pytest

from conftest import bits_equal, mismatch_report
from presets_util import to_oracle_params

import slime_mold_b200 as sm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_wgsl_golden as mk  # noqa: E402
import wgsl_reference as wr  # noqa: E402

SMALL_TAGS = [c[0] for c in mk.SMALL_CASES]                   # 1x1, 2x3, 13x7, 37x5 maps
CASE_TAGS = [c[0] for c in mk.CASES] + SMALL_TAGS
LOCKSTEP_TAGS = [c[0] for c in mk.CASES if c[2].get("pheromone_deposition_amount", 1.0) >= 1.0] + SMALL_TAGS


def load(tag):
    return np.load(os.path.join(GOLD, f"wgsl_{tag}.npz"))


def uniform_of(g):
    return sm.SimSizeUniform.from_buffer_copy(g["params"].tobytes())


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_oracle_sequential_equals_shader(oracle, tag):
    g = load(tag)
    u = uniform_of(g)
    sim = oracle.Sim(to_oracle_params(oracle, u), g["agents0"], g["trail0"])
    for k in range(1, int(g["frames"]) + 1):
        sim.step_sequential(1, inplace_diffuse=True)
        assert bits_equal(sim.agents, g[f"seq_agents{k}"]), mismatch_report(sim.agents, g[f"seq_agents{k}"], f"agents, frame {k}")
        assert bits_equal(sim.trail, g[f"seq_trail{k}"]), mismatch_report(sim.trail, g[f"seq_trail{k}"], f"trail, frame {k}")


@pytest.mark.parametrize("tag", LOCKSTEP_TAGS)
def test_oracle_phase_split_equals_shader_lockstep(oracle, tag):
    g = load(tag)
    u = uniform_of(g)
    sim = oracle.Sim(to_oracle_params(oracle, u), g["agents0"], g["trail0"])
    for k in range(1, int(g["frames"]) + 1):
        sim.step(1)
        assert bits_equal(sim.agents, g[f"lock_agents{k}"]), mismatch_report(sim.agents, g[f"lock_agents{k}"], f"agents, frame {k}")
        assert bits_equal(sim.trail, g[f"lock_trail{k}"]), mismatch_report(sim.trail, g[f"lock_trail{k}"], f"trail, frame {k}")


def test_oracle_low_deposit_differs_only_where_cells_are_shared(oracle):
    """dep < 1: lockstep loses all but one of the deposits into a shared cell, phase_split adds them all (DESIGN.md
    section 2) -- agents are identical, the trail differs, and only upwards."""
    g = load("lowdep")
    u = uniform_of(g)
    sim = oracle.Sim(to_oracle_params(oracle, u), g["agents0"], g["trail0"])
    sim.step(1)
    assert bits_equal(sim.agents, g["lock_agents1"])
    assert (sim.trail >= g["lock_trail1"]).all() and (sim.trail > g["lock_trail1"]).any()


def test_oracle_edge_agents_equal_shader(oracle):
    g = load("edge")
    u = uniform_of(g)
    p = to_oracle_params(oracle, u)
    sim = oracle.Sim(p, g["agents0"], g["trail0"])
    oracle.agents_sequential(sim.agents, sim.trail, p)
    assert bits_equal(sim.agents, g["seq_agents1"]), mismatch_report(sim.agents, g["seq_agents1"], "agents")
    assert bits_equal(sim.trail, g["seq_trail1"])
    sim = oracle.Sim(p, g["agents0"], g["trail0"])
    oracle.agents_phase_split(sim.agents, sim.trail, sim.counts, p)
    oracle.deposit_merge(sim.trail, sim.counts, p.pheromone_deposition_amount)
    assert bits_equal(sim.agents, g["lock_agents1"]), mismatch_report(sim.agents, g["lock_agents1"], "agents")
    assert bits_equal(sim.trail, g["lock_trail1"])


def test_oracle_display_equals_shader(oracle):
    g = load("display")
    for tw, th in g["shapes"]:
        ref = g[f"rgba_{tw}x{th}"]
        got = oracle.display(g["trail"], g["lut"], int(tw), int(th))
        assert np.array_equal(got, ref), f"{tw}x{th}: {np.count_nonzero(got != ref)} bytes differ"


@pytest.mark.parametrize("tag", LOCKSTEP_TAGS)
def test_oracle_within_tolerance_of_libm_backend(oracle, tag):
    """Same shader, numpy's libm sin / cos instead of the spec's: speeds identical; positions within 2e-3 px; headings
    within 1e-5 + 0.011 * jitter for > 97 % of the agents (a 1-ulp sin difference is multiplied by 43758 in the hash)."""
    g = load(tag)
    u = uniform_of(g)
    a, b = g["lock_agents1"], g["libm_agents1"]
    assert np.array_equal(a[:, 3], b[:, 3])
    dang = np.abs(a[:, 2] - b[:, 2])
    dang = np.minimum(dang, np.abs(dang - np.float32(2 * np.pi)))
    tol = 1e-5 + 0.011 * u.agent_jitter
    ok = dang <= tol
    assert ok.mean() > (0.97 if len(ok) > 100 else 0.9)
    move = u.agent_speed_max * 0.016
    dx = np.abs(a[ok, 0] - b[ok, 0])
    dy = np.abs(a[ok, 1] - b[ok, 1])
    dx = np.minimum(dx, np.abs(dx - u.width))
    dy = np.minimum(dy, np.abs(dy - u.height))
    assert max(dx.max(), dy.max()) <= 2e-3 + move * tol


# ---- only where the reference tree exists (the build container): the fixtures ARE outputs of the shader file ----
needs_reference = pytest.mark.skipif(not wr.have_reference(), reason="/root/reference is not present on this machine")


@needs_reference
def test_fixture_digests_match_the_shader_files():
    cd, dd = wr.source_digest(wr.shader_source("compute.wgsl")), wr.source_digest(wr.shader_source("display.wgsl"))
    for tag in CASE_TAGS + ["edge"]:
        assert str(load(tag)["shader_sha256"]) == cd, tag
    assert str(load("display")["shader_sha256"]) == dd


@needs_reference
@pytest.mark.parametrize("tag", ["default", "mesh_jitter", "tiny_1x1", "ragged_13x7"])
def test_shader_rerun_reproduces_fixture(oracle, tag):
    g = load(tag)
    u = uniform_of(g)
    src = wr.shader_source("compute.wgsl")
    sim = wr.ShaderSim(src, u, g["agents0"], g["trail0"], wr.SpecMath(oracle))
    sim.frame("lockstep")
    assert bits_equal(sim.agents.data, g["lock_agents1"]) and bits_equal(sim.trail2d, g["lock_trail1"])
    sim = wr.ShaderSim(src, u, g["agents0"], g["trail0"], wr.SpecMath(oracle))
    sim.run_agents("sequential")
    sim.run_decay()
    assert bits_equal(sim.agents.data, g["seq_agents1"])


@needs_reference
def test_display_shader_rerun_reproduces_fixture():
    g = load("display")
    u = uniform_of(g)
    tw, th = 17, 13
    got = wr.run_display(wr.shader_source("display.wgsl"), u, g["trail"], g["lut"], tw, th)
    assert np.array_equal(got, g[f"rgba_{tw}x{th}"])


def test_interpreter_typing_rules():
    """The interpreter itself: WGSL's literal typing, truncating integer ops, conversions, schedules."""
    from wgsl_interp import F32, I32, Interpreter, StorageArray
    src = """
    @group(0) @binding(0) var<storage, read_write> buf: array<f32>;
    const K: f32 = 0.1;
    fn helper(a: f32) -> f32 { if (a > 1.0) { return a * 2.0; } return 0.5; }
    @compute @workgroup_size(4)
    fn main(@builtin(global_invocation_id) id: vec3<u32>) {
        let i = i32(id.x);
        if (i >= i32(arrayLength(&buf))) { return; }
        var acc = 0.0;
        for (var k = -1; k <= 1; k++) { acc += f32((i + k + 4) % 4); }
        let left = buf[(i + 3) % 4];
        buf[i] = left + acc * K + helper(f32(i)) + f32(-7 / 2) + f32(-7 % 3) + f32(i32(-2.7)) + f32(u32(3.9));
    }
    """
    def run(schedule):
        it = Interpreter(src)
        b = StorageArray(np.array([1, 2, 3, 4], np.float32))
        it.bind("buf", b)
        it.dispatch("main", [(i, 0, 0) for i in range(8)], schedule)
        return b.data
    f = np.float32
    const = f(-3.0) + f(-1.0) + f(-2.0) + f(3.0)               # -7/2 = -3, -7%3 = -1, i32(-2.7) = -2, u32(3.9) = 3

    def expect(left, i):
        acc = f(sum((i + k + 4) % 4 for k in (-1, 0, 1)))
        h = f(i) * f(2.0) if i > 1 else f(0.5)
        v = f(left) + acc * f(0.1)
        for t in (h, f(-3.0), f(-1.0), f(-2.0), f(3.0)):
            v = f(v + t)
        return v
    lock = run("lockstep")
    init = [1, 2, 3, 4]
    assert [expect(init[(i + 3) % 4], i) for i in range(4)] == list(lock)
    seq = run("sequential")
    cur = [f(v) for v in init]
    for i in range(4):
        cur[i] = expect(cur[(i + 3) % 4], i)
    assert cur == list(seq) and list(seq) != list(lock)
    assert const == f(-3.0) and isinstance(I32(1), np.int32) and F32 is np.float32


@pytest.mark.parametrize("tag", LOCKSTEP_TAGS + ["edge"])
def test_engine_arithmetic_on_host_equals_shader(oracle, hostcheck, tag):
    """The product's __host__ __device__ statements (agent_core.cuh / trail_core.cuh, instantiated for the host by
    tests/hostcheck) against the shader's lockstep frames -- the CPU-side rehearsal of tests/test_gpu_wgsl.py."""
    import ctypes as C

    def P(a, t):
        return a.ctypes.data_as(C.POINTER(t))
    g = load(tag)
    u = uniform_of(g)
    p = to_oracle_params(oracle, u)
    H, W = g["trail0"].shape
    a = np.ascontiguousarray(g["agents0"]).copy()
    tr = np.ascontiguousarray(g["trail0"]).copy()
    cn = np.zeros((H, W), np.uint32)
    out = np.empty_like(tr)
    frames = int(g["frames"]) if "frames" in g.files else 1
    for k in range(1, frames + 1):
        hostcheck.hc_agents_phase_split(P(a, C.c_float), None, C.c_uint64(a.shape[0]), P(tr, C.c_float), P(cn, C.c_uint32), C.byref(p))
        assert bits_equal(a, g[f"lock_agents{k}"]), mismatch_report(a, g[f"lock_agents{k}"], f"agents, frame {k}")
        if tag == "edge":
            break
        hostcheck.hc_trail_pass(P(tr, C.c_float), P(cn, C.c_uint32), P(out, C.c_float), C.byref(p))
        tr, out = out, tr
        assert bits_equal(tr, g[f"lock_trail{k}"]), mismatch_report(tr, g[f"lock_trail{k}"], f"trail, frame {k}")
