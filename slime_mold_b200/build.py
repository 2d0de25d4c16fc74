"""Build recipe: nvcc -> slime_mold_b200/libslime_b200.so (sm_100a only, in-tree).

    python -m slime_mold_b200.build [--force]

Flags that matter for parity with the arithmetic spec: --fmad=false (no
FMA contraction; every fused multiply-add in the kernels is an explicit __fmaf_rn),
default -prec-div/-prec-sqrt/-ftz=false.  -lineinfo so ncu's source page maps to
the .cu files.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libslime_b200.so")
SOURCES = ["engine.cu", "gauss.cu", "exchange.cu"]
HEADERS = ["engine.h", "kernels.cuh", "agent_core.cuh", "trail_core.cuh", "gauss_stream.cuh", "gauss_rows.cuh", "device_math.cuh",
           os.path.join("..", "..", "include", "slime_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.sep not in p or os.path.exists(p)):
            return p
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    # the translation units are compiled side by side (the kernel templates make each of them a minute or two of ptxas)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc_path(), *NVCC_FLAGS, "-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed = "", False
    for cmd, pr in procs:
        out, _ = pr.communicate()
        log += " ".join(cmd) + "\n" + out
        failed |= pr.returncode != 0
    if not failed:
        cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        failed = res.returncode != 0
    for obj in objs:
        if os.path.exists(obj):
            os.remove(obj)
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed (see slime_mold_b200/build.log)")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
