"""Build libalaya_b200.so (nvcc, sm_100a) and the _alayalitepy pybind11 module (g++) in-tree.

    python -m alayalite_b200.build [--force]

nvcc cross-compiles without a GPU; the shared objects stay next to the sources (git-ignored) so
that they travel to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libalaya_b200.so")
PYMOD = os.path.join(HERE, "_alayalitepy" + sysconfig.get_config_var("EXT_SUFFIX"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CU_SOURCES = ["index.cu", "search.cu", "search_sq8.cu", "search_sq4.cu", "quant.cu", "bruteforce.cu", "bruteforce_tc.cu", "io.cu", "build.cu", "update.cu", "group.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "alaya_b200.h"))
    return hs


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + cmd[0])
    return r.stdout + r.stderr


def build_lib(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = _headers()
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s + ".o")
        if force or not _newer(obj, [src] + hdrs):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([NVCC] + NVCC_FLAGS + extra + ["-c", src, "-o", obj])
    out = ""
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            out = "".join(ex.map(_run, jobs))
    objs = [os.path.join(OBJ, s + ".o") for s in srcs]
    if force or jobs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    return out


def build_pymod(force=False):
    src = os.path.join(CSRC, "pybind_module.cpp")
    if not force and _newer(PYMOD, [src, LIB] + _headers()):
        return
    import pybind11

    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(HERE, "..", "include"),
           src, "-o", PYMOD, "-L" + HERE, "-lalaya_b200", "-Wl,-rpath,$ORIGIN"]
    _run(cmd)


def build(force=False, verbose=False):
    out = build_lib(force, verbose)
    build_pymod(force)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
