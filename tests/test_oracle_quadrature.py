"""The CPU oracle of the trajectory consumers (oracle/quad_oracle.hpp: hermiteInterpolate, cumtrapz, cumsimpson and
their function variants) pinned on the reference's own test cases, tests/test_integrate.nim:5-22, 67-95:
f(x) = a cos x with a = 2, X = linspace(0, 3pi/2, 17), Y = f(X), expected 2 sin x within the reference's tolerances;
plus the behaviours the reference's code implies but its tests never reach (restated from the code: sorting, duplicate
trimming, the sorted / unsorted branches of hermiteInterpolate, even- and odd-length Simpson)."""
import math

import numpy as np
import pytest

import oracle as O

X_END = 3.0 / 2.0 * math.pi
X = O.linspace(0.0, X_END, 17)                       # tests/test_integrate.nim:20
Y = np.array([2.0 * math.cos(x) for x in X])         # :21
CUM_Y = np.array([2.0 * math.sin(x) for x in X])     # :22


def is_close(a, b, tol):  # utils.nim:474-479 for floats
    return abs(a - b) <= tol


def test_reference_cumtrapz_discrete():  # tests/test_integrate.nim:67-70
    v = O.cumtrapz(Y, X, scalar=True)
    assert len(v) == 17 and all(is_close(v[i], CUM_Y[i], 1e-1) for i in range(17))
    assert v[0] == 0.0


def test_reference_cumsimpson_discrete():  # :82-85
    v = O.cumsimpson(Y, X, scalar=True)
    assert len(v) == 17 and all(is_close(v[i], CUM_Y[i], 1e-3) for i in range(17))


@pytest.mark.parametrize("dx,tol_t,tol_s", [(0.1, 1e-1, 1e-3), (1e-3, 1e-1, 1e-3)])
def test_reference_function_variants(dx, tol_t, tol_s):  # :77-80, :92-95 (and the default-dx cases with a coarser dx)
    f = lambda t: 2.0 * math.cos(t)
    v, evals = O.cumtrapz_fn(f, X, dx=dx, scalar=True)
    assert len(v) == 17 and all(is_close(v[i], CUM_Y[i], tol_t) for i in range(17))
    assert evals >= (X_END + 1.0) / dx  # integrates to max(X) + 1.0 (integrate.nim:160)
    v, evals = O.cumsimpson_fn(f, X, dx=dx, scalar=True)
    assert len(v) == 17 and all(is_close(v[i], CUM_Y[i], tol_s) for i in range(17))
    assert evals == round(X_END / dx) + 2  # linspace(min, max, toInt((max-min)/dx) + 2), integrate.nim:396


def test_vector_instantiation_equals_scalar_per_component():
    """T = Vector runs the same arithmetic per component as T = float (the reference's generics)."""
    Yv = np.stack([Y, 0.5 * Y, -Y], axis=1)
    for fn in (O.cumtrapz, O.cumsimpson):
        v = fn(Yv, X)
        for c, s in enumerate((1.0, 0.5, -1.0)):
            ref = fn(s * Y, X, scalar=True)
            assert np.array_equal(v[:, c].view(np.uint64), ref.view(np.uint64))


def test_sorting_and_duplicate_rule():  # utils.nim:360-420
    Xd = np.array([0.0, 1.0, 0.5, 1.0, 2.0])
    Yd = np.array([1.0, 3.0, 2.0, 3.0, 5.0])
    assert O.cumtrapz(Yd, Xd, scalar=True).tolist() == [0.0, 0.75, 2.0, 6.0]  # 4 distinct x, sorted
    with pytest.raises(ValueError, match="impure y-duplicates"):
        O.cumtrapz(np.array([1.0, 3.0, 2.0, 4.0, 5.0]), Xd, scalar=True)
    # cumsimpson interpolates back onto the ORIGINAL X (unsorted, with its duplicate): 5 values again
    v = O.cumsimpson(Yd, Xd, scalar=True)
    assert len(v) == 5 and v[1] == v[3] and v[0] == 0.0
    with pytest.raises(ValueError, match="at least 3 elements"):
        O.cumsimpson(Yd[:2], Xd[:2], scalar=True)
    with pytest.raises(O.Defect):
        O.cumtrapz(np.array([]), np.array([]), scalar=True)


@pytest.mark.parametrize("m", [3, 4, 5, 8, 9])
def test_simpson_is_exact_for_cubics_on_uneven_grids(m):
    """Simpson's rule on two unequal intervals integrates quadratics exactly and the Hermite interpolation of the
    running integral reproduces a quartic with exact slopes only approximately, so test at the knots: even points."""
    rng = np.random.default_rng(m)
    x = np.sort(rng.uniform(0.0, 2.0, m))
    f = lambda t: 1.0 + 2.0 * t + 3.0 * t * t
    F = lambda t: t + t * t + t ** 3
    v = O.cumsimpson(f(x), x, scalar=True)
    knots = list(range(0, m if m % 2 else m - 1, 2)) + ([m - 1] if m % 2 == 0 else [])
    for k in knots:
        assert abs(v[k] - (F(x[k]) - F(x[0]))) <= 1e-12 * max(1.0, abs(F(x[k]))), (m, k)


def test_hermite_interpolate_branches():  # utils.nim:287-312
    t = np.array([0.0, 1.0, 2.0, 4.0])
    y = t ** 3
    dy = 3 * t ** 2
    xs = np.array([0.0, 0.25, 1.0, 3.0, 4.0])
    v = O.hermite_interpolate(xs, t, y, dy, scalar=True)
    assert np.allclose(v, xs ** 3, rtol=1e-14, atol=1e-14) and v[-1] == 64.0  # cubic reproduced; x == t[high] copies y[high]
    # sorted x: samples before t[0] stall the scan (nothing after them is produced), samples beyond the end are dropped
    assert len(O.hermite_interpolate([-1.0, 0.5], t, y, dy, scalar=True)) == 0
    assert len(O.hermite_interpolate([0.5, 5.0], t, y, dy, scalar=True)) == 1
    assert len(O.hermite_interpolate([3.5, 4.0, 4.0], t, y, dy, scalar=True)) == 2  # the end point is appended once
    # unsorted x: every sample is searched; outside -> ValueError
    v = O.hermite_interpolate([3.0, 0.25, 4.0, 1.0], t, y, dy, scalar=True)
    assert np.allclose(v, np.array([3.0, 0.25, 4.0, 1.0]) ** 3, rtol=1e-14)
    with pytest.raises(ValueError, match="not in interval"):
        O.hermite_interpolate([3.0, 0.25, 5.0], t, y, dy, scalar=True)


def test_simpson_weights_reduce_to_the_classic_rule():
    a, b, e = O.simpson_weights(0.5, 0.5)
    assert (a, b, e) == (0.5 / 3.0, 4.0 * 0.5 / 3.0, 0.5 / 3.0) or np.allclose([a, b, e], [1 / 6, 4 / 6, 1 / 6], rtol=1e-15)
    a, b, e = O.simpson_weights(0.5, 0.5, tail=True)  # last interval of an even-length set: (5, 8, -1) h / 12
    assert np.allclose([a, b, e], [5 * 0.5 / 12, 8 * 0.5 / 12, -0.5 / 12], rtol=1e-15)


# ---------------------------------------------------------------------------------------------------
# two independent restatements of the Nim source agree bit for bit (C++ oracle vs tests/pyref_quad.py)
# ---------------------------------------------------------------------------------------------------
import pyref_quad as P  # noqa: E402


def _bits(a):
    return [float(v).hex() for v in a]


@pytest.mark.parametrize("seed", range(30))
def test_oracle_equals_python_restatement(seed):
    rng = np.random.default_rng(1000 + seed)
    m = int(rng.integers(3, 14))
    X = rng.uniform(0.0, 3.0, m)
    Y = rng.uniform(-2.0, 2.0, m)
    if seed % 3 == 0:
        X = np.sort(X)
    if seed % 4 == 0:  # pure duplicates
        X[1], Y[1] = X[0], Y[0]
        if m > 4:
            X[4], Y[4] = X[2], Y[2]
    try:
        want_t, want_s = P.cumtrapz(Y.tolist(), X.tolist()), None
    except ValueError:
        with pytest.raises(ValueError):
            O.cumtrapz(Y, X, scalar=True)
        return
    assert _bits(O.cumtrapz(Y, X, scalar=True)) == _bits(want_t)
    try:
        want_s = P.cumsimpson(Y.tolist(), X.tolist())
    except ValueError:
        with pytest.raises(ValueError):
            O.cumsimpson(Y, X, scalar=True)
        return
    assert _bits(O.cumsimpson(Y, X, scalar=True)) == _bits(want_s)
    t = np.sort(rng.uniform(0.0, 3.0, m))
    y, dy = rng.uniform(-1, 1, m), rng.uniform(-1, 1, m)
    x = rng.uniform(t[0], t[-1], 7)
    x[0] = t[-1]
    if seed % 2:
        x = np.sort(x)
    assert _bits(O.hermite_interpolate(x, t, y, dy, scalar=True)) == _bits(P.hermite_interpolate(x.tolist(), t.tolist(), y.tolist(), dy.tolist()))


@pytest.mark.parametrize("dx", [0.1, 0.0371])
def test_oracle_function_variants_equal_python_restatement(dx):
    f = lambda t: 2.0 * math.cos(t) + 0.25 * t
    for Xs in (list(X), [X[3], X[0], X[16], X[7], X[7], X[1]]):
        v, _ = O.cumtrapz_fn(f, Xs, dx=dx, scalar=True)
        assert _bits(v) == _bits(P.cumtrapz_fn(f, Xs, dx))
        v, _ = O.cumsimpson_fn(f, Xs, dx=dx, scalar=True)
        assert _bits(v) == _bits(P.cumsimpson_fn(f, Xs, dx))


def test_hermite_spline_restatements_agree():
    """utils.nim:273-279 is written with `^`; the oracle's product form (u*u, t*(t*t)) must give the same bits."""
    rng = np.random.default_rng(4)
    for _ in range(2000):
        x1, w = rng.uniform(-2, 2), 10.0 ** rng.uniform(-3, 1)
        x2 = x1 + w
        x = x1 + w * rng.uniform(0, 1)
        y1, y2, d1, d2 = rng.uniform(-3, 3, 4)
        a = O.hermite(x, x1, x2, [y1], [y2], [d1], [d2])[0]
        assert float(a).hex() == float(P.hermite_spline(x, x1, x2, y1, y2, d1, d2)).hex()
