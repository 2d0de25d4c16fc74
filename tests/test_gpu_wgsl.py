"""The CUDA engine (through the C ABI) against outputs of the REFERENCE'S OWN SHADER SOURCE -- tests/golden/wgsl_*.npz,
produced from /root/reference/src/compute.wgsl and display.wgsl by tests/golden/make_wgsl_golden.py (lockstep schedule:
every load of a dispatch sees the buffers as they were when the dispatch started).  Bit for bit.  -m gpu.
Nothing here touches the oracle or /root/reference: the committed vectors are the checker."""
import os

import numpy as np
import pytest

import slime_mold_b200 as sm
from conftest import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOCKSTEP_TAGS = ["default", "sponge", "waves", "snake", "mesh_jitter", "tiny_1x1", "tiny_2x3", "ragged_13x7", "ragged_37x5"]


def load(tag):
    return np.load(os.path.join(GOLD, f"wgsl_{tag}.npz"))


def backend_for(g, **kw):
    u = sm.SimSizeUniform.from_buffer_copy(g["params"].tobytes())
    be = sm.CudaBackend.new(int(u.width), int(u.height), sm.Settings.default(), agent_count=g["agents0"].shape[0], device=0, **kw)
    be.write_uniform(u)
    be.write_agents(g["agents0"])
    be.write_trail(g["trail0"])
    return be


@pytest.mark.parametrize("sort_interval", [0, 1])
@pytest.mark.parametrize("tag", LOCKSTEP_TAGS)
def test_engine_equals_shader_frames(engine_lib, tag, sort_interval):
    g = load(tag)
    with backend_for(g, sort_interval=sort_interval) as be:
        for k in range(1, int(g["frames"]) + 1):
            be.step(1)
            a, t = be.read_agents(), be.read_trail()
            assert bits_equal(a, g[f"lock_agents{k}"]), f"{tag} frame {k}: " + mismatch_report(a, g[f"lock_agents{k}"], "agents")
            assert bits_equal(t, g[f"lock_trail{k}"]), f"{tag} frame {k}: " + mismatch_report(t, g[f"lock_trail{k}"], "trail")


def test_engine_equals_shader_agents_when_deposits_are_fractional(engine_lib):
    """dep < 1 (u32 count path): the agents still match the shader bit for bit; the trail is the order-free sum of all
    deposits where the racing shader keeps one, so it can only be higher."""
    g = load("lowdep")
    with backend_for(g) as be:
        be.step(1)
        a, t = be.read_agents(), be.read_trail()
    assert bits_equal(a, g["lock_agents1"]), mismatch_report(a, g["lock_agents1"], "agents")
    assert (t >= g["lock_trail1"]).all()


def test_engine_equals_shader_on_edge_agents(engine_lib):
    """NaN / inf / huge headings, positions outside the map and on the seams: one `main` dispatch of the shader.  The
    engine's step also runs the trail pass, so the agents are compared (the deposits are covered by the frame tests)."""
    g = load("edge")
    with backend_for(g) as be:
        be.step(1)
        a = be.read_agents()
    assert bits_equal(a, g["lock_agents1"]), mismatch_report(a, g["lock_agents1"], "agents")


def test_engine_display_equals_shader(engine_lib):
    g = load("display")
    u = sm.SimSizeUniform.from_buffer_copy(g["params"].tobytes())
    with sm.CudaBackend.new(int(u.width), int(u.height), sm.Settings.default(), agent_count=16, device=0) as be:
        be.write_trail(g["trail"])
        be.set_lut(g["lut"])
        for tw, th in g["shapes"]:
            frame = be.render(int(tw), int(th))
            ref = g[f"rgba_{tw}x{th}"]
            assert np.array_equal(frame, ref), f"{tw}x{th}: {np.count_nonzero(frame != ref)} bytes differ"
