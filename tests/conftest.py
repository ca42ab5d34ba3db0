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


def pytest_addoption(parser):
    parser.addoption("--host-emulation", action="store_true", default=False,
                     help="TEST INFRASTRUCTURE: run the `gpu` tests on the CPU against tests/host_emul/_build/libb200rk_emul.so — the product's "
                          "host sources and kernels compiled by g++ with every kernel launch emulated thread by thread (no GPU, no NVRTC)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    if config.getoption("--host-emulation"):
        # Only the test-suite can do this: the product package has no switch for it and raises without a CUDA device.
        sys.path.insert(0, os.path.join(ROOT, "tests", "host_emul"))
        import build_emul_lib
        from numericalnim_b200 import _capi
        _capi.LIB_PATH = build_emul_lib.build(sanitize=os.environ.get("B200RK_TEST_EMULATION_SANITIZE", ""))
        os.environ["B200RK_TEST_HOST_EMULATION"] = "1"


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
