"""rust/ cannot be compiled here (no rustc): keep it mechanically consistent with include/slime_b200.h -- every C
function declared in the header is declared in rust/src/ffi.rs with the same number of arguments, every #[repr(C)]
struct lists the header's fields in the header's order with types of the same width, and the backend only calls
functions the FFI declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "slime_b200.h")).read()
HEADER = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
FFI = re.sub(r"//[^\n]*", "", open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read())
BACKEND = re.sub(r"//[^\n]*", "", open(os.path.join(ROOT, "rust", "src", "cuda_backend.rs")).read())

C_WIDTH = {"uint32_t": 4, "int32_t": 4, "float": 4, "double": 8, "uint64_t": 8}
RS_WIDTH = {"u32": 4, "i32": 4, "f32": 4, "f64": 8, "u64": 8}


def nargs(arglist):
    arglist = arglist.strip()
    return 0 if arglist in ("", "void") else arglist.count(",") + 1


def c_functions():
    return {m.group(1): nargs(m.group(2)) for m in re.finditer(r"\b(sm_[a-z_0-9]+)\s*\(([^()]*)\)\s*;", HEADER)}


def rust_functions():
    return {m.group(1): nargs(m.group(2)) for m in re.finditer(r"pub fn (sm_[a-z_0-9]+)\s*\(([^()]*)\)", FFI)}


def test_every_header_function_is_declared_with_the_same_arity():
    c, r = c_functions(), rust_functions()
    assert len(c) >= 33
    assert c == r, {k: (c.get(k), r.get(k)) for k in set(c) | set(r) if c.get(k) != r.get(k)}


def c_struct(name):
    body = re.search(r"typedef struct " + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", HEADER, re.S).group(1)
    out = []
    for ty, names in re.findall(r"(uint32_t|int32_t|uint64_t|float|double|sm_tuning)\s+([^;]+);", body):
        for n in names.split(","):
            n = n.strip()
            arr = re.match(r"(\w+)\[(\d+)\]$", n)
            if ty == "sm_tuning":
                out.append((n, sum(w for _, w in c_struct("sm_tuning"))))
            elif arr:
                out.append((arr.group(1), C_WIDTH[ty] * int(arr.group(2))))
            else:
                out.append((n, C_WIDTH[ty]))
    return out


def rust_struct(name):
    body = re.search(r"pub struct " + name + r"\s*\{(.*?)\}", FFI, re.S).group(1)
    out = []
    for n, t in re.findall(r"pub (\w+):\s*([^,\n]+),", body):
        t = t.strip()
        arr = re.match(r"\[(\w+);\s*(\d+)\]$", t)
        if t == "sm_tuning":
            out.append((n, sum(w for _, w in rust_struct("sm_tuning"))))
        elif arr:
            out.append((n, RS_WIDTH[arr.group(1)] * int(arr.group(2))))
        else:
            out.append((n, RS_WIDTH[t]))
    return out


def test_repr_c_structs_mirror_the_header():
    for name in ("sm_params", "sm_tuning", "sm_config", "sm_timing", "sm_trail_stats"):
        assert c_struct(name) == rust_struct(name), name
        assert re.search(r"#\[repr\(C\)\]\s*(#\[derive[^\]]*\]\s*)?pub struct " + name, FFI), name
    assert sum(w for _, w in c_struct("sm_params")) == 56          # SimSizeUniform, src/main.rs:29-46


def test_status_codes_and_flags_match():
    for name, val in re.findall(r"(SM_(?:OK|ERR_[A-Z_]+))\s*=\s*(-?\d+)", HEADER):
        assert re.search(r"pub const " + name + r": c_int = " + val + ";", FFI), name
    assert "SM_FLAG_GAUSSIAN_BLUR: u32 = 1 << 0" in FFI and "SM_FLAG_NO_SORT: u32 = 1 << 1" in FFI
    assert "SM_COMM_ID_BYTES: usize = 128" in FFI and "#define SM_COMM_ID_BYTES 128" in HEADER


def test_backend_calls_only_declared_functions():
    declared = set(rust_functions())
    used = set(re.findall(r"\b(sm_[a-z_0-9]+)\s*\(", BACKEND))
    assert used and used <= declared, used - declared
    # the reference-facing surface INTEGRATION.md section 1 maps onto the ABI
    for method in ("new", "write_uniform", "init_agents", "write_agents", "read_agents", "reassign_agent_speeds",
                   "set_agent_count", "clear_trail", "resize", "step", "set_lut", "render", "read_trail"):
        assert re.search(r"pub fn " + method + r"\b", BACKEND), method


# ---- module paths: every `use` of the shim must resolve against the reference's crate layout --------------------------
REF_SRC = "/root/reference/src"


def _use_paths(src):
    out = []
    for m in re.finditer(r"^use\s+([^;]+);", src, re.M):
        path = m.group(1).strip()
        g = re.match(r"(.*)::\{(.*)\}$", path, re.S)
        if g:
            out += [g.group(1) + "::" + leaf.strip() for leaf in g.group(2).split(",") if leaf.strip()]
        else:
            out.append(path)
    return out


def test_use_paths_resolve_against_the_reference_crates():
    """The reference is a lib crate (src/lib.rs: `pub mod lut_manager; pub mod presets; pub mod settings;`) plus a bin crate
    (src/main.rs) that imports it as `slime_mold::..` and owns `SimSizeUniform`.  The shim's files are modules of the BIN
    crate, so: `slime_mold::<m>::<Item>` needs `pub mod <m>` in lib.rs and a `pub` item in src/<m>.rs; `crate::<Item>` needs
    the item in main.rs; `crate::ffi::*` needs rust/src/ffi.rs (added to main.rs as `mod ffi;`)."""
    import pytest
    if not os.path.isdir(REF_SRC):
        pytest.skip("the reference checkout is not on this machine")
    lib_rs = open(os.path.join(REF_SRC, "lib.rs")).read()
    main_rs = open(os.path.join(REF_SRC, "main.rs")).read()
    cargo = open(os.path.join(os.path.dirname(REF_SRC), "Cargo.toml")).read()
    crate_name = re.search(r'^name\s*=\s*"([^"]+)"', cargo, re.M).group(1).replace("-", "_")
    assert crate_name == "slime_mold"
    own_modules = {"ffi", "cuda_backend"}
    checked = 0
    for path in _use_paths(BACKEND) + _use_paths(FFI):
        parts = path.split("::")
        if parts[0] == "std":
            continue
        if parts[0] == crate_name:                                   # lib crate
            mod, item = parts[1], parts[2]
            assert re.search(r"pub mod " + mod + r"\s*;", lib_rs), f"{path}: lib.rs has no `pub mod {mod}`"
            src = open(os.path.join(REF_SRC, mod + ".rs")).read()
            assert re.search(r"pub (struct|enum|fn|type|const) " + item + r"\b", src), f"{path}: no pub item {item} in {mod}.rs"
        elif parts[0] == "crate":                                    # bin crate (main.rs is the crate root)
            if parts[1] in own_modules:
                assert os.path.exists(os.path.join(ROOT, "rust", "src", parts[1] + ".rs")), path
            else:
                assert len(parts) == 2, f"{path}: nested bin-crate paths are not expected"
                assert re.search(r"^(pub )?(struct|enum|fn|type|const) " + parts[1] + r"\b", main_rs, re.M), \
                    f"{path}: main.rs defines no {parts[1]}"
        else:
            raise AssertionError(f"{path}: neither std, the lib crate nor the bin crate")
        checked += 1
    assert checked >= 4
    # the methods the shim calls on reference types exist there
    assert "agent_count" in open(os.path.join(REF_SRC, "settings.rs")).read()
    for field in re.findall(r"lut\.(\w+)", BACKEND):
        assert re.search(r"pub " + field + r"\b", open(os.path.join(REF_SRC, "lut_manager.rs")).read()), field
