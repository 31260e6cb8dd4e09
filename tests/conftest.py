"""Shared fixtures. `-m gpu` tests need a B200; everything else runs on the CPU-only build box."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


@pytest.fixture(scope="session")
def orc():
    """Our CPU restatement of the reference algorithm (oracle/restate.cpp)."""
    from oracle import oracle as O

    if not os.path.exists(O.ORC_SO):
        O.build(ref=os.path.isdir("/root/reference/include"))
    return O.Checker("orc")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled in place (oracle/_ref). Prebuilt on the GPU box."""
    from oracle import oracle as O

    if not O.have_ref():
        if os.path.isdir("/root/reference/include"):
            O.build(ref=True)
        else:
            pytest.skip("oracle/_ref/libref_oracle.so not available")
    return O.Checker("ref")


@pytest.fixture(scope="session")
def cabi():
    from alayalite_b200 import _cabi

    _cabi.lib()
    return _cabi


def recall(res, gt):
    hits = sum(len(set(map(int, r)) & set(map(int, g))) for r, g in zip(res, gt))
    return hits / float(res.shape[0] * res.shape[1])


@pytest.fixture(scope="session")
def small_sets():
    """Seeded datasets shared by many tests: name -> (base, queries)."""
    from alayalite_b200.synthetic import integer_grid, low_rank, uniform

    return {
        "lowrank": low_rank(4000, 48, 200, latent=8, seed=3),
        "lowrank128": low_rank(6000, 128, 200, latent=12, seed=4),
        "grid": integer_grid(3000, 32, 200, seed=7, hi=12),
        "uniform": uniform(1000, 128, 10, seed=0),
        "odd": low_rank(1500, 37, 50, latent=6, seed=5),
    }
