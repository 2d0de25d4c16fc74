"""Generates tests/golden/*.npz and reference_presets.json.

    python tests/golden/make_golden.py

* golden_<preset>.npz : seeded runs of the CPU oracle (phase_split) -- state after 1 and
  after 25 steps on a 96x64 map with 2500 agents, plus a diffusion-only pass.  The
  reference itself cannot run here (no rustc / Vulkan), so these are ORACLE outputs:
  they pin the oracle against drift and give the GPU tests committed vectors, they do
  not pin the oracle to the reference -- tests/golden/make_wgsl_golden.py does that (outputs of the
  reference's shader source).
* reference_presets.json : the preset / default values parsed out of
  /root/reference/src/presets.rs and settings.rs (only when /root/reference exists),
  used to check the Python/C++ mirrors of the reference's configuration surface.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import slime_oracle as so  # noqa: E402
from presets_util import PRESET_NAMES, preset_uniform, random_trail, to_oracle_params  # noqa: E402

W, H, N, SEED = 96, 64, 2500, 2024


def make_preset_golden(name):
    u = preset_uniform(name, W, H)
    p = to_oracle_params(so, u)
    ag0 = so.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, SEED)
    sim = so.Sim(p, ag0)
    sim.step(1)
    a1, t1 = sim.agents.copy(), sim.trail.copy()
    sim.step(24)
    a25, t25 = sim.agents.copy(), sim.trail.copy()
    field = random_trail(W, H, seed=7, density=0.5)
    d1 = so.trail_pass(field, p, counts=None)
    fn = os.path.join(HERE, "golden_" + name.lower().replace(" ", "_") + ".npz")
    np.savez_compressed(fn, params=np.frombuffer(bytes(u), dtype=np.uint8), agents0=ag0, agents1=a1, trail1=t1,
                        agents25=a25, trail25=t25, field=field, diffused=d1, seed=SEED)
    return fn


def parse_reference():
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        return None
    out = {"defaults": {}, "presets": {}}
    txt = open(os.path.join(ref, "settings.rs")).read()
    for m in re.finditer(r"pub const (\w+): (?:f32|usize|u32) = ([0-9_.]+);", txt):
        out["defaults"][m.group(1)] = float(m.group(2).replace("_", ""))
    ptxt = open(os.path.join(ref, "presets.rs")).read()
    for m in re.finditer(r'Preset::new\(\s*"([^"]+)"\.to_string\(\),\s*Settings \{(.*?)\.\.Settings::default\(\)', ptxt, re.S):
        fields = {k: float(v.replace("_", "")) for k, v in re.findall(r"(\w+): ([0-9_.]+),", m.group(2))}
        out["presets"][m.group(1)] = fields
    out["presets"]["Default"] = {}
    return out


if __name__ == "__main__":
    so.build()
    for n in PRESET_NAMES:
        print(make_preset_golden(n))
    ref = parse_reference()
    if ref is not None:
        with open(os.path.join(HERE, "reference_presets.json"), "w") as f:
            json.dump(ref, f, indent=1, sort_keys=True)
        print("reference_presets.json", len(ref["presets"]), "presets")
