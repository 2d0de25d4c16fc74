"""The C++ host mirror (slime_mold_b200/host/slime_backend.hpp): same names / defaults / packing as the
reference's Rust surface, and the same bytes as the Python mirror."""
import os
import subprocess

import numpy as np
import pytest

import slime_mold_b200 as sm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_cpp", "test_host_mirror.cpp")
EXE = os.path.join(ROOT, "tests", "host_cpp", "test_host_mirror")


@pytest.fixture(scope="module")
def exe(engine_lib):
    hdr = os.path.join(ROOT, "slime_mold_b200", "host", "slime_backend.hpp")
    if not os.path.exists(EXE) or max(os.path.getmtime(SRC), os.path.getmtime(hdr)) > os.path.getmtime(EXE):
        cxx = "/usr/bin/g++-13" if os.path.exists("/usr/bin/g++-13") else "g++"
        libdir = os.path.join(ROOT, "slime_mold_b200")
        subprocess.run([cxx, "-std=c++17", "-O1", "-o", EXE, SRC, "-L" + libdir, "-lslime_b200", "-Wl,-rpath," + libdir],
                       check=True)
    return EXE


def test_cpp_presets_pack_the_same_bytes(exe):
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    lines = dict(l.split(":", 1) for l in out.strip().split("\n"))
    pm = sm.init_preset_manager()
    assert list(lines) == pm.get_preset_names()
    for name in pm.get_preset_names():
        s = pm.get_preset(name).settings
        u = sm.SimSizeUniform.new(1920, 1080, s.pheromone_decay_factor, s)
        assert lines[name] == bytes(u).hex(), name


def test_cpp_fails_loudly_without_device(exe):
    if sm.device_count() > 0:
        pytest.skip("a B200 is present")
    out = subprocess.run([exe, "nodevice"], capture_output=True, text=True, check=True).stdout
    assert "nodevice:-5" in out and "no CPU fallback" in out


@pytest.mark.gpu
def test_cpp_backend_matches_oracle(exe, oracle):
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, check=True).stdout
    got = [l for l in out.split("\n") if l.startswith("gpu_checksum:")][0].split(":")[1]
    assert "error_check:ok -1" in out
    s = sm.init_preset_manager().get_preset("Waves").settings
    from presets_util import preset_uniform, to_oracle_params
    sim = oracle.Sim(to_oracle_params(oracle, preset_uniform("Waves", 256, 128)),
                     oracle.init_agents(20000, 256, 128, s.agent_speed_min, s.agent_speed_max, 5))
    sim.step(10)
    h = 1469598103934665603
    for arr in (sim.agents, sim.trail):
        for b in np.ascontiguousarray(arr).view(np.uint8).ravel().tolist():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert got == f"{h:016x}"
