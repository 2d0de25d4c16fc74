"""The Python mirror of the reference's configuration surface (settings.rs, presets.rs,
SimSizeUniform of main.rs:29-67) against values parsed out of the reference
(tests/golden/reference_presets.json, written by make_golden.py from /root/reference)."""
import ctypes as C
import json
import os
import struct

import slime_mold_b200 as sm
from slime_mold_b200 import settings as st

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_presets.json")


def test_defaults_match_reference_constants():
    ref = json.load(open(GOLD))["defaults"]
    for k, v in ref.items():
        name = {"DEFAULT_WIDTH": "DEFAULT_WIDTH", "DEFAULT_HEIGHT": "DEFAULT_HEIGHT"}.get(k, k)
        assert float(getattr(st, name)) == v, k
    s = sm.Settings.default()
    assert (s.agent_count, s.window_width, s.window_height) == (10_000_000, 1600, 900)
    assert s.agent_possible_starting_headings == (0.0, 360.0)


def test_presets_match_reference():
    ref = json.load(open(GOLD))["presets"]
    pm = sm.init_preset_manager()
    assert pm.get_preset_names() == ["Default", "Sponge", "Firecracker Trees", "Threads", "Curls", "Waves", "Snake", "Mesh"]
    assert set(ref) == set(pm.get_preset_names())
    d = sm.Settings.default()
    for name, fields in ref.items():
        s = pm.get_preset(name).settings
        for f in d.__dataclass_fields__:
            expect = fields.get(f, getattr(d, f))
            assert getattr(s, f) == expect, (name, f)
    assert pm.get_preset("nope") is None


def test_uniform_packing_is_repr_c():
    s = sm.Settings.default().clone(agent_jitter=0.25, blur_radius=3.0, blur_sigma=1.5)
    u = sm.SimSizeUniform.new(1920, 1080, s.pheromone_decay_factor, s)
    raw = bytes(u)
    assert len(raw) == 56
    vals = struct.unpack("<II11fI", raw)
    assert vals[0:2] == (1920, 1080)
    exp = [10.0, 0.25, 30.0, 50.0, 0.43, 0.3, 20.0, 1.0, 1.0, 3.0, 1.5]
    assert [round(v, 6) for v in vals[2:13]] == [round(C.c_float(e).value, 6) for e in exp]
    assert vals[13] == 0
