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
    for ty, names in re.findall(r"(uint32_t|int32_t|uint64_t|float|double)\s+([^;]+);", body):
        out += [(n.strip(), C_WIDTH[ty]) for n in names.split(",")]
    return out


def rust_struct(name):
    body = re.search(r"pub struct " + name + r"\s*\{(.*?)\}", FFI, re.S).group(1)
    return [(n, RS_WIDTH[t]) for n, t in re.findall(r"pub (\w+):\s*(\w+),", body)]


def test_repr_c_structs_mirror_the_header():
    for name in ("sm_params", "sm_config", "sm_timing", "sm_trail_stats"):
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
