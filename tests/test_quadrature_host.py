"""Host-side decisions of the trajectory consumers (csrc/quadrature.cu), checked on the CPU against the oracle: which
interval every sample of hermiteInterpolate uses and the spline's scalar factors (b200rk_hermite_plan), and Simpson's
coefficient triples (b200rk_simpson_weights). With these pinned, what is left to the GPU tests is element-wise
arithmetic. No compute calls, no device."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from numericalnim_b200 import _capi


def plan(x, t):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    t = np.ascontiguousarray(np.asarray(t, dtype=np.float64))
    j = np.zeros(max(x.size, 1), dtype=np.int32)
    k = np.zeros(max(x.size, 1), dtype=np.int32)
    f = np.zeros(4 * max(x.size, 1))
    n = C.c_size_t(0)
    _capi.check(_capi.lib().b200rk_hermite_plan(x.ctypes.data, x.size, t.ctypes.data, t.size, j.ctypes.data, k.ctypes.data, f.ctypes.data, C.byref(n)))
    return j[: n.value], k[: n.value], f[: 4 * n.value].reshape(-1, 4)


def apply_plan(j, k, f, y, dy):
    """The kernel's arithmetic (quad_kernels.cuh: hermite_elem) in Python floats: same operations, same order."""
    out = []
    for o in range(len(j)):
        if k[o] == 1:
            out.append(y[j[o]])
        else:
            a = j[o]
            out.append(((y[a] * f[o, 0] + dy[a] * f[o, 1]) + y[a + 1] * f[o, 2]) + dy[a + 1] * f[o, 3])
    return np.array(out)


def pyref_plan(x, t):
    """utils.nim:287-312 restated independently of both the oracle and the library (interval indices only)."""
    res, xi, th, xh = [], 0, len(t) - 1, len(x) - 1
    if all(x[i] <= x[i + 1] for i in range(len(x) - 1)):
        for i in range(0, th):
            while t[i] <= x[xi] and x[xi] < t[i + 1]:
                res.append((i, 0))
                xi += 1
                if xh < xi:
                    break
            if xh < xi:
                break
        if x[xh] == t[th]:
            res.append((th, 1))
    else:
        for a in x:
            for i in range(0, th):
                if t[i] <= a and a < t[i + 1]:
                    res.append((i, 0))
                    break
            else:
                if a == t[th]:
                    res.append((th, 1))
                else:
                    raise ValueError("not in interval")
    return res


@pytest.mark.parametrize("seed", range(240))
def test_plan_matches_oracle_and_restatement(seed):
    rng = np.random.default_rng(seed)
    nt = int(rng.integers(1, 9))
    t = np.sort(rng.uniform(0.0, 4.0, nt))
    if seed % 5 == 0 and nt > 2:
        t[1] = t[2]  # a zero-length interval can never be chosen
    if seed % 7 == 3 and nt > 2:
        t = rng.permutation(t)  # unsorted data: the literal left-to-right scan of the reference decides
    nx = int(rng.integers(1, 12))
    pool = np.concatenate([t, rng.uniform(t.min() - (0.5 if seed % 3 == 0 else 0.0), t.max() + (0.5 if seed % 4 == 0 else 0.0), 8)])
    x = rng.choice(pool, nx)
    if seed % 2 == 0:
        x = np.sort(x)
    y, dy = rng.uniform(-1, 1, nt), rng.uniform(-1, 1, nt)
    try:
        want = pyref_plan(list(x), list(t))
    except ValueError:
        with pytest.raises(ValueError, match="not in interval"):
            plan(x, t)
        with pytest.raises(ValueError, match="not in interval"):
            O.hermite_interpolate(x, t, y, dy, scalar=True)
        return
    j, k, f = plan(x, t)
    assert list(zip(j.tolist(), k.tolist())) == want
    ref = O.hermite_interpolate(x, t, y, dy, scalar=True)
    got = apply_plan(j, k, f, y, dy)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (x, t)


def test_plan_quirks_and_errors():
    t = [0.0, 1.0, 2.0, 4.0]
    assert len(plan([-1.0, 0.5], t)[0]) == 0               # a sorted sample before t[0] stalls the scan (utils.nim:291-298)
    assert plan([0.5, 5.0], t)[0].tolist() == [0]            # samples past the end are dropped
    j, k, _ = plan([3.5, 4.0, 4.0], t)
    assert list(zip(j.tolist(), k.tolist())) == [(2, 0), (3, 1)]  # the end point is appended once (utils.nim:299-300)
    assert plan([4.0], [4.0])[1].tolist() == [1]             # a data set of one point
    with pytest.raises(ValueError):
        plan([], t)
    with pytest.raises(ValueError):
        plan([1.0], [])


@pytest.mark.parametrize("tail", [0, 1])
def test_simpson_weights_match_oracle_bitwise(tail):
    rng = np.random.default_rng(5 + tail)
    L = _capi.lib()
    for _ in range(500):
        h1, h2 = (float(v) for v in 10.0 ** rng.uniform(-6, 2, 2))
        a, b, e = C.c_double(0), C.c_double(0), C.c_double(0)
        assert L.b200rk_simpson_weights(tail, h1, h2, C.byref(a), C.byref(b), C.byref(e)) == 0
        assert (a.value.hex(), b.value.hex(), e.value.hex()) == tuple(v.hex() for v in O.simpson_weights(h1, h2, tail=bool(tail)))
    # equal intervals: the classic (1, 4, 1) h/3 and, for the tail, (5, 8, -1) h/12
    a, b, e = C.c_double(0), C.c_double(0), C.c_double(0)
    L.b200rk_simpson_weights(0, 0.5, 0.5, C.byref(a), C.byref(b), C.byref(e))
    assert np.allclose([a.value, b.value, e.value], [0.5 / 3, 4 * 0.5 / 3, 0.5 / 3], rtol=1e-15)
    L.b200rk_simpson_weights(1, 0.5, 0.5, C.byref(a), C.byref(b), C.byref(e))
    assert np.allclose([a.value, b.value, e.value], [5 * 0.5 / 12, 8 * 0.5 / 12, -0.5 / 12], rtol=1e-15)
