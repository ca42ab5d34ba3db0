"""Trajectory consumers on the GPU (csrc/quadrature.cu + quad_kernels.cuh) vs the CPU oracle: hermiteInterpolate
(utils.nim:282-312), cumtrapz / cumsimpson on sampled trajectories (integrate.nim:119-135, 330-378) and their function
variants (integrate.nim:138-175, 379-400), T = device vector. All arithmetic is element-wise, so every result is
compared BIT FOR BIT (the kernels keep the reference's association, no FMA); the host-side scalars (sorting,
duplicate rule, interval search, Simpson's coefficients, spline factors) are the reference's expressions."""
import math

import numpy as np
import pytest
from conftest import assert_bitwise_equal

import oracle as O

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 4, 5, 7, 1023, 4096, 65536 + 7]


@pytest.fixture(scope="module")
def nn():
    import numericalnim_b200 as nn
    nn.default_context()
    return nn


def _traj(rng, m, n, special=False):
    Y = rng.uniform(-2.0, 2.0, (m, n))
    if special and n >= 4:
        Y[0, 0] = np.inf       # "the right kind of zero" y0 - y0 is NaN there, like in the reference
        Y[1, 1] = -0.0
        Y[min(2, m - 1), 2] = 5e-324
        Y[m - 1, 3] = np.nan
    return Y


def assert_same(got, ref, what):
    """Bit for bit; a NaN must be a NaN in the same place (the sign / payload of a generated NaN is not IEEE-specified
    and differs between x86 and the GPU, as in test_gpu_parity.py::test_stage_accum_special_values)."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    ng, nr = np.isnan(got), np.isnan(ref)
    assert np.array_equal(ng, nr), what + ": NaN pattern differs"
    assert_bitwise_equal(got[~nr], ref[~nr], what)


def _dev(nn, A):
    return [nn.newVector(row) for row in A]


def _host(vs):
    return np.array([v.to_numpy() for v in vs]) if vs else np.empty((0, 0))


@pytest.mark.parametrize("vec_width", [4, 2])
@pytest.mark.parametrize("m", [1, 2, 3, 6, 17, 40])
def test_cumtrapz_bitwise(nn, m, vec_width):
    ctx = nn.default_context()
    rng = np.random.default_rng(100 + m)
    try:
        ctx.set("vec_width", vec_width)
        for n in SIZES:
            X = np.sort(rng.uniform(0.0, 3.0, m))
            Y = _traj(rng, m, n, special=(n == 1023 and m >= 3))
            got = _host(nn.cumtrapz(_dev(nn, Y), X))
            assert_same(got, O.cumtrapz(Y, X), f"cumtrapz m={m} n={n}")
    finally:
        ctx.set("vec_width", 4)


@pytest.mark.parametrize("vec_width", [4, 2])
@pytest.mark.parametrize("m", [3, 4, 5, 8, 17, 40])
def test_cumsimpson_bitwise(nn, m, vec_width):
    """Odd and even lengths (the even case ends with the one-interval tail rule, integrate.nim:361-373)."""
    ctx = nn.default_context()
    rng = np.random.default_rng(200 + m)
    try:
        ctx.set("vec_width", vec_width)
        for n in SIZES:
            X = np.sort(rng.uniform(0.0, 3.0, m))
            Y = _traj(rng, m, n, special=(n == 1023))
            got = _host(nn.cumsimpson(_dev(nn, Y), X))
            assert_same(got, O.cumsimpson(Y, X), f"cumsimpson m={m} n={n}")
    finally:
        ctx.set("vec_width", 4)


def test_unsorted_input_and_duplicates(nn):
    """X unsorted with pure duplicates: sorted + trimmed for cumtrapz (fewer results), interpolated back onto the
    original X for cumsimpson (as many results as inputs); impure duplicates raise ValueError."""
    rng = np.random.default_rng(7)
    n = 4099
    base = rng.uniform(-1.0, 1.0, (6, n))
    X = np.array([0.0, 1.0, 0.5, 1.0, 2.0, 1.5, 0.5, 3.0])
    Y = np.stack([base[0], base[1], base[2], base[1], base[3], base[4], base[2], base[5]])
    dv = _dev(nn, Y)
    got = _host(nn.cumtrapz(dv, X))
    assert got.shape[0] == 6
    assert_bitwise_equal(got, O.cumtrapz(Y, X), "cumtrapz unsorted + duplicates")
    got = _host(nn.cumsimpson(dv, X))
    assert got.shape[0] == 8
    assert_bitwise_equal(got, O.cumsimpson(Y, X), "cumsimpson unsorted + duplicates")
    Ybad = Y.copy()
    Ybad[3, n - 1] += 1e-9  # same x = 1.0, one component differs
    for fn in (nn.cumtrapz, nn.cumsimpson):
        with pytest.raises(ValueError, match="impure y-duplicates"):
            fn(_dev(nn, Ybad), X)
    Ynan = Y.copy()
    Ynan[1, 5] = Ynan[3, 5] = np.nan  # NaN != NaN: even "identical" duplicates are impure (utils.nim:371-373)
    with pytest.raises(ValueError, match="impure y-duplicates"):
        nn.cumtrapz(_dev(nn, Ynan), X)
    with pytest.raises(ValueError, match="impure y-duplicates"):
        O.cumtrapz(Ynan, X)


def test_quadrature_errors(nn):
    a, b = nn.newVector(np.ones(8)), nn.newVector(np.ones(9))
    with pytest.raises(ValueError, match="same size"):
        nn.cumtrapz([a, b], [0.0, 1.0])
    with pytest.raises(ValueError, match="at least 3 elements"):
        nn.cumsimpson([a, a], [0.0, 1.0])
    with pytest.raises(ValueError, match="at least 3 elements"):
        nn.cumsimpson([a, a, a], [0.0, 1.0, 1.0])  # 2 distinct points after trimming
    with pytest.raises(ValueError, match="same length"):
        nn.cumtrapz([a, a], [0.0, 1.0, 2.0])
    with pytest.raises(ValueError):
        nn.cumtrapz([], [])
    with pytest.raises(ValueError, match="NaN"):
        nn.cumtrapz([a, a], [0.0, float("nan")])


@pytest.mark.parametrize("n", SIZES)
def test_hermite_interpolate_bitwise(nn, n):
    rng = np.random.default_rng(300 + n)
    t = np.array([0.0, 0.5, 1.25, 2.0, 4.0])
    y, dy = rng.uniform(-1, 1, (5, n)), rng.uniform(-1, 1, (5, n))
    gy, gdy = _dev(nn, y), _dev(nn, dy)
    # sorted: several samples per interval, skipped intervals, both end points, the end point asked twice
    xs = np.array([0.0, 0.1, 0.4, 0.5, 1.3, 1.9, 3.0, 3.5, 4.0, 4.0])
    got = _host(nn.hermiteInterpolate(xs, t, gy, gdy))
    ref = O.hermite_interpolate(xs, t, y, dy)
    assert got.shape[0] == ref.shape[0] == 9
    assert_bitwise_equal(got, ref, "hermiteInterpolate sorted")
    # unsorted: every sample searched on its own
    xu = np.array([3.0, 0.1, 4.0, 1.3, 0.0, 3.5, 0.4])
    assert_bitwise_equal(_host(nn.hermiteInterpolate(xu, t, gy, gdy)), O.hermite_interpolate(xu, t, y, dy), "hermiteInterpolate unsorted")
    # the reference's quirks: a sorted sample before t[0] stalls the scan; samples past the end are dropped
    assert len(nn.hermiteInterpolate([-1.0, 0.5], t, gy, gdy)) == len(O.hermite_interpolate([-1.0, 0.5], t, y, dy)) == 0
    assert len(nn.hermiteInterpolate([0.5, 5.0], t, gy, gdy)) == 1
    with pytest.raises(ValueError, match="not in interval"):
        nn.hermiteInterpolate([3.0, 0.25, 5.0], t, gy, gdy)


def test_dense_output_of_a_solve_is_hermite_interpolate(nn):
    """The two ends of the path meet: solveODE's dense output (ode.nim:512-524) at the requested times equals
    hermiteInterpolate over the accepted steps — here checked through the public API on an RK4 solve, where the
    accepted steps are known (fixed dt), by re-interpolating the trajectory sampled at every step."""
    n = 1000
    lam = np.linspace(0.1, 3.0, n)
    y0 = np.linspace(1.0, 2.0, n)
    glam = nn.newVector(lam)
    rhs = nn.rhsDiagLinear(glam)
    steps = [0.01 * k for k in range(0, 51)]
    o = nn.newODEoptions(dt=0.01)
    t_all, y_all = nn.solveODE(rhs, nn.newVector(y0), steps, o, integrator="rk4")
    ref = O.solve_vector("rk4", O.rhs_diag_linear(lam), y0, steps, O.new_options(dt=0.01))
    got = _host(y_all)
    assert_bitwise_equal(got, ref.y, "rk4 trajectory")
    dy_all = [-(glam.hmul(v)) for v in y_all]
    xs = [0.005, 0.0151, 0.2, 0.33333, 0.4999]
    out = _host(nn.hermiteInterpolate(xs, t_all, y_all, dy_all))
    exp = O.hermite_interpolate(xs, np.array(t_all), ref.y, -(lam * ref.y))
    assert_bitwise_equal(out, exp, "hermiteInterpolate over a solved trajectory")
    exact = y0 * np.exp(-np.outer(xs, lam))
    assert np.max(np.abs(out - exact)) < 1e-7


@pytest.mark.parametrize("dx", [0.1, 0.013])
def test_function_variants_bitwise(nn, dx):
    """cumtrapz(f, X, ctx, dx) streams (4 vectors alive); cumsimpson(f, X, ctx, dx) composes the discrete rule and the
    interpolation like the reference. f(x, ctx) = cos(x) * ctx["a"] as in tests/test_integrate.nim:6."""
    n = 4099
    a = np.linspace(1.0, 3.0, n)
    ga = nn.newVector(a)
    X = O.linspace(0.0, 1.5 * math.pi, 17)
    numctx = nn.newNumContext()
    numctx["a"] = ga
    f = lambda x, ctx: math.cos(x) * ctx["a"]          # scalar * GpuVector: one kernel
    f_host = lambda x: a * math.cos(x)
    for name, ofn in (("cumtrapz", O.cumtrapz_fn), ("cumsimpson", O.cumsimpson_fn)):
        ref, evals = ofn(f_host, X, dx=dx, n=n)
        got = _host(getattr(nn, name)(f, X, ctx=numctx, dx=dx, like=ga))
        assert_bitwise_equal(got, ref, f"{name}(f, X, dx={dx})")
        assert np.max(np.abs(got - np.outer(np.sin(X), a))) < (1e-1 if name == "cumtrapz" else 1e-3) * 3.0  # test_integrate.nim:77-95
    # unsorted X goes through the per-sample search; a sample outside [min, max + 1] cannot occur by construction
    Xu = X[[3, 0, 16, 7, 7, 1]]
    ref, _ = O.cumtrapz_fn(f_host, Xu, dx=dx, n=n)
    assert_bitwise_equal(_host(nn.cumtrapz(f, Xu, ctx=numctx, dx=dx, like=ga)), ref, "cumtrapz(f) unsorted X")
    ref, _ = O.cumsimpson_fn(f_host, Xu, dx=dx, n=n)
    assert_bitwise_equal(_host(nn.cumsimpson(f, Xu, ctx=numctx, dx=dx, like=ga)), ref, "cumsimpson(f) unsorted X")


def test_function_variant_counts_and_errors(nn):
    n = 64
    ga = nn.newVector(np.ones(n))
    calls = []

    def f(x, ctx):
        calls.append(x)
        return math.cos(x) * ga

    X = [0.0, 0.5, 1.0]
    nn.cumtrapz(f, X, dx=0.25, like=ga)
    ref, evals = O.cumtrapz_fn(lambda x: np.ones(n) * math.cos(x), X, dx=0.25, n=n)
    assert len(calls) == evals  # min(X) .. max(X) + 1.0 in steps of dx, no extra evaluation
    calls.clear()
    nn.cumsimpson(f, X, dx=0.25, like=ga)
    assert len(calls) == round(1.0 / 0.25) + 2
    with pytest.raises(ZeroDivisionError):
        nn.cumtrapz(lambda x, ctx: (1 / 0) * ga, X, dx=0.25, like=ga)  # an exception in the integrand reaches the caller
    with pytest.raises(ValueError, match="dx must be"):
        nn.cumtrapz(f, X, dx=0.0, like=ga)


def test_full_size_cumtrapz_property(nn):
    """BASELINE size (2^23 per vector): too big for the oracle's allocating Vector ops in seconds, so check
    size-independent properties: linearity in the data for a constant trajectory (integral = c * (x - x0) exactly
    for power-of-two spacings) and agreement of cumtrapz's last value with Simpson's on a linear-in-time trajectory."""
    n = 1 << 23
    m = 9
    X = [0.25 * k for k in range(m)]
    c = nn.newVector(np.full(n, 3.0))
    out = nn.cumtrapz([c] * m, X)
    assert len(out) == m
    for k in (0, 4, 8):
        v = out[k].to_numpy()
        assert v[0] == v[n // 2] == v[n - 1] == 3.0 * X[k]
    base = nn.newVector(np.linspace(-1.0, 1.0, n))
    lin = [base * (1.0 + 0.5 * k) for k in range(m)]     # y(x) = base * (1 + 2x): both rules are exact
    t_last = nn.cumtrapz(lin, X)[-1].to_numpy()
    s_last = nn.cumsimpson(lin, X)[-1].to_numpy()
    exact = np.linspace(-1.0, 1.0, n) * (2.0 + 2.0 * 2.0)  # int_0^2 (1 + 2x) dx = 6
    assert np.max(np.abs(t_last - exact)) <= 1e-14 * 6 and np.max(np.abs(s_last - exact)) <= 1e-14 * 6
