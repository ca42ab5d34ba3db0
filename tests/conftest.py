import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

# Compiled user right-hand sides (csrc/jit.cu) are cached in-tree so that a test run on a fresh GPU box reuses what
# scripts/jit_prewarm.py compiled on the CPU box (NVRTC needs no GPU); keyed by NVRTC version + kernels.cuh contents.
os.environ.setdefault("B200RK_JIT_CACHE", os.path.join(ROOT, ".jitcache"))

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


@pytest.fixture(scope="session")
def golden_trajectories():
    with open(os.path.join(GOLDEN, "trajectories.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_tableaux():
    with open(os.path.join(GOLDEN, "tableaux.json")) as fh:
        raw = json.load(fh)
    return {m: {k: float.fromhex(v) for k, v in d.items()} for m, d in raw.items()}


def bits(a):
    """View float64 array as uint64 so comparisons are bit-exact (distinguishes -0.0, NaN payloads)."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_bitwise_equal(a, b, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    ne = bits(a) != bits(b)
    if ne.any():
        idx = np.flatnonzero(ne.ravel())[:5]
        raise AssertionError(f"{what}: {ne.sum()} of {a.size} elements differ bitwise; first at {idx}: "
                             f"{a.ravel()[idx]} vs {b.ravel()[idx]}")
