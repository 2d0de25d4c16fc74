import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _gpu_present() -> bool:
    return any(os.path.exists(p) for p in ("/dev/nvidia0", "/dev/nvidiactl"))


def pytest_collection_modifyitems(config, items):
    if _gpu_present():
        return
    skip = pytest.mark.skip(reason="no NVIDIA device node on this machine")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand with oracle/Makefile."""
    from oracle import slime_oracle as so
    so.build()
    so.lib()
    return so


@pytest.fixture(scope="session")
def engine_lib():
    """libslime_b200.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from slime_mold_b200 import build as b
    b.build()
    from slime_mold_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def hostcheck():
    """TEST-ONLY host instantiation of the engine's __host__ __device__ arithmetic."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "hostcheck")
    so_path = os.path.join(d, "libhostcheck.so")
    src = os.path.join(d, "hostcheck.cpp")
    csrc = os.path.join(ROOT, "slime_mold_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("device_math.cuh", "agent_core.cuh", "trail_core.cuh", "gauss_stream.cuh", "gauss_rows.cuh")]
    if not os.path.exists(so_path) or any(os.path.getmtime(p) > os.path.getmtime(so_path) for p in deps):
        cxx = "/usr/bin/g++-13" if os.path.exists("/usr/bin/g++-13") else "g++"
        subprocess.run([cxx, "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC",
                        "-shared", "-pthread", "-o", so_path, src], check=True)
    return C.CDLL(so_path)


def bits_equal(a, b) -> bool:
    """Bit-exact comparison of two f32 arrays, treating any NaN as equal to any NaN."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def mismatch_report(a, b, name="array") -> str:
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    bad = a.view(np.uint32) != b.view(np.uint32)
    bad &= ~(np.isnan(a) & np.isnan(b))
    n = int(bad.sum())
    if n == 0:
        return f"{name}: identical"
    idx = np.argwhere(bad)[:5]
    ex = [(tuple(i), float(a[tuple(i)]), float(b[tuple(i)])) for i in idx]
    return f"{name}: {n}/{a.size} elements differ, first: {ex}"


@pytest.fixture(scope="session")
def eq():
    return bits_equal
