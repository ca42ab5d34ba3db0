"""Paths that are OFF by default because they were written after the round's GPU budget was spent: their kernels are
verified bit for bit by host emulation (tests/test_kernel_host_emulation.py) but have not run on a B200 yet. The tests
below compare them with the default (GPU-verified) path; they are non-strict xfail so that their first GPU run is
reported (XPASS / xfail) without being able to turn the suite red or to mask anything (the file runs last)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="experimental, off by default: verified by host emulation, first GPU run pending")]


@pytest.fixture(scope="module")
def nn():
    import numericalnim_b200 as nn
    nn.default_context()
    return nn


@pytest.mark.parametrize("vec_width", [4, 2])
def test_cumsimpson_single_pass_equals_two_kernel_path(nn, vec_width):
    """knob fuse_simpson=1: Simpson scan + Hermite interpolation in ONE kernel (simpson_fused_kernel), bit-identical to
    simpson_scan_kernel + hermite_many_kernel, one launch instead of two."""
    ctx = nn.default_context()
    rng = np.random.default_rng(5)
    try:
        ctx.set("vec_width", vec_width)
        for m in (3, 4, 5, 8, 17, 40):
            for n in (1, 3, 4, 5, 1023, 65536 + 7):
                X = np.sort(rng.uniform(0.0, 3.0, m))
                if m == 8:  # unsorted samples with a pure duplicate: results come back in the caller's order
                    X = np.concatenate([X[::-1], X[::-1][2:3]])
                Y = rng.uniform(-2.0, 2.0, (m, n))
                if m == 8:
                    Y = np.concatenate([Y[::-1], Y[::-1][2:3]])
                dv = [nn.newVector(r) for r in Y]
                res = {}
                for fuse in (0, 1):
                    ctx.set("fuse_simpson", fuse)
                    l0 = ctx.stats()["launches"]
                    out = nn.cumsimpson(dv, X)
                    res[fuse] = (np.array([v.to_numpy() for v in out]), ctx.stats()["launches"] - l0)
                assert res[0][0].shape == res[1][0].shape
                assert np.array_equal(res[0][0].view(np.uint64), res[1][0].view(np.uint64)), (m, n)
                if m != 8:
                    assert res[1][1] == 1 and res[0][1] == 2
    finally:
        ctx.set("fuse_simpson", 0)
        ctx.set("vec_width", 4)
