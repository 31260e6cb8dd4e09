"""The C++ plugin surface (include/alaya_b200.hpp): compiles, satisfies the reference's concepts, and on a
GPU gives the same answers as the C ABI it wraps."""
import os
import subprocess
import sysconfig

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
LIBDIR = os.path.join(ROOT, "alayalite_b200")
REF = "/root/reference"


def _compile(src, out, std="c++17", extra=()):
    cmd = ["g++", f"-std={std}", "-O1", "-w", "-I", os.path.join(ROOT, "include"), "-I", CPP, *extra,
           os.path.join(CPP, src), "-o", out, "-L", LIBDIR, "-lalaya_b200", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_header_instantiates_and_refuses_without_gpu(tmp_path):
    exe = str(tmp_path / "syntax")
    _compile("surface_syntax.cpp", exe)
    import torch
    r = subprocess.run([exe], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 42, "constructing a DeviceSpace without a CUDA device must throw"
    else:
        assert r.returncode == 0


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include")), reason="reference checkout not present")
def test_classes_satisfy_reference_concepts(tmp_path):
    site = sysconfig.get_paths()["purelib"]
    extra = ["-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(REF, "include"),
             "-I", os.path.join(site, "flashinfer/data/spdlog/include"),
             "-I", os.path.join(site, "tilelang/3rdparty/tvm/3rdparty/dmlc-core/include")]
    _compile("surface_concepts.cpp", str(tmp_path / "concepts"), std="c++20", extra=extra)


@pytest.mark.gpu
def test_cpp_surface_end_to_end(tmp_path):
    exe = str(tmp_path / "gpu")
    _compile("surface_gpu.cpp", exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    head, ids = r.stdout.strip().splitlines()[:2]
    assert head == "same=1 solo_same=1 sorted=1 dist_ok=1 n=4000", head
    ids = np.array(ids.split(), dtype=np.uint32).reshape(50, 10)
    assert len(np.unique(ids[0])) == 10 and ids.max() < 4000


@pytest.mark.gpu
def test_plain_c_host_runs_a_two_shard_search(tmp_path):
    """ab_group_* from a C program: no torch, no NCCL, no C++."""
    exe = str(tmp_path / "shard")
    lib_dir = os.path.join(ROOT, "alayalite_b200")
    cmd = ["gcc", "-O2", "-std=c11", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "shard_gpu.c"), "-o", exe,
           "-L", lib_dir, "-lalaya_b200", "-Wl,-rpath," + lib_dir]
    c = subprocess.run(cmd, capture_output=True, text=True)
    assert c.returncode == 0, c.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    fields = dict(f.split("=") for f in r.stdout.split())
    assert float(fields["recall"]) >= 0.97 and fields["sorted"] == "1" and int(fields["ids_from_second_shard"]) > 0
