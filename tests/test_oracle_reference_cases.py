"""The reference's own tests for the hot path, restated against the CPU oracle (config 1 of
BASELINE.json: "plumbing, no GPU"). Sources: tests/test_ode.nim, tests/test_vector.nim,
tests/test_utils.nim of the reference (cited per test)."""
import math

import numpy as np
import pytest

import oracle as O

TSPAN = None


def tspan():
    global TSPAN
    if TSPAN is None:
        TSPAN = O.linspace(-10.0, 10.0, 100)  # tests/test_ode.nim:15
    return TSPAN


def correct(t):
    return np.array([math.exp(-0.1 * x) for x in t])  # tests/test_ode.nim:8


OO = dict(relTol=1e-8, dt=1e-6)  # tests/test_ode.nim:9
OOV = dict(relTol=1e-8, dt=1e-2)  # tests/test_ode.nim:10

# (integrator, options kwargs or None, tol) — scalar cases, tests/test_ode.nim:24-136
SCALAR_CASES = [
    ("dopri54", None, 1e-4), ("dopri54", OO, 1e-8), ("rk4", None, 1e-4),
    ("heun2", None, 1e-10), ("heun2", None, 1e-8), ("ralston2", None, 1e-10), ("kutta3", None, 1e-10),
    ("heun3", None, 1e-10), ("ralston3", None, 1e-10), ("ssprk3", None, 1e-10), ("ralston4", None, 1e-10),
    ("kutta4", None, 1e-10), ("rk21", None, 1e-6), ("bs32", None, 1e-6), ("tsit54", None, 1e-4),
    ("tsit54", OO, 1e-8), ("vern65", None, 1e-4), ("vern65", OO, 1e-8),
]


@pytest.mark.parametrize("integrator,opts,tol", SCALAR_CASES)
def test_ode_scalar(integrator, opts, tol):
    o = O.new_options(**opts) if opts else None
    t, y, _ = O.solve_scalar(integrator, 1.0, tspan(), o)
    assert np.array_equal(t, tspan())  # `check t == tspan`
    ref = correct(t)
    assert len(y) == len(t)
    for v, c in zip(y, ref):
        assert O.is_close(float(v), float(c), tol)


def test_ode_scalar_rk4_dt_1e_6_prefix():
    """tests/test_ode.nim:42-46 ("RK4, dt = 1e-6") is 2x10^7 steps; run the same options over the first two
    output intervals on either side of 0 (same code path, bounded time)."""
    ts = tspan()[48:52]
    t, y, _ = O.solve_scalar("rk4", 1.0, ts, O.new_options(**OO))
    assert np.array_equal(t, ts)
    for v, c in zip(y, correct(t)):
        assert O.is_close(float(v), float(c), 1e-8)


# Vector cases, tests/test_ode.nim:139-197
VECTOR_CASES = [
    ("dopri54", None, 1e-4), ("dopri54", OOV, 1e-8), ("rk4", None, 1e-4), ("rk4", OOV, 1e-8),
    ("heun2", None, 1e-8), ("heun2", OOV, 1e-5), ("tsit54", None, 1e-4), ("tsit54", OOV, 1e-8),
    ("vern65", None, 1e-4), ("vern65", OOV, 1e-8),
]


@pytest.mark.parametrize("integrator,opts,tol", VECTOR_CASES)
def test_ode_vector3(integrator, opts, tol):
    o = O.new_options(**opts) if opts else None
    sol = O.solve_vector(integrator, O.rhs_scale(-0.1), [1.0, 1.0, 1.0], tspan(), o)
    assert np.array_equal(sol.t, tspan())
    assert sol.y.shape == (100, 3)
    for row, c in zip(sol.y, correct(sol.t)):
        assert O.is_close(row, np.array([c, c, c]), tol)


def test_unknown_integrator_is_value_error():  # ode.nim:650-651
    with pytest.raises(ValueError):
        O.solve_scalar("rk5", 1.0, tspan())


def test_integrator_name_is_case_insensitive():  # ode.nim:607
    t1, y1, _ = O.solve_scalar("DOPRI54", 1.0, tspan())
    t2, y2, _ = O.solve_scalar("dopri54", 1.0, tspan())
    assert np.array_equal(y1, y2)


def test_options_validation():  # ode.nim:95-102
    with pytest.raises(ValueError):
        O.new_options(dtMax=1e-5, dtMin=1e-4)
    with pytest.raises(ValueError):
        O.new_options(scaleMax=0.5)
    with pytest.raises(ValueError):
        O.new_options(scaleMin=1.5)
    o = O.new_options(dt=-1e-3, absTol=-1e-5, dtMax=-1.0, dtMin=-1e-6)
    assert (o.dt, o.absTol, o.dtMax, o.dtMin) == (1e-3, 1e-5, 1.0, 1e-6)


# ---- tests/test_vector.nim known-answers for the operators on the path -----------------------
def test_vector_add_sub():  # test_vector.nim:26-30, 47-52
    v1, v2 = [1.1, 2.2, 3.3], [3.3, 2.2, 1.0]
    assert O.vector_binop(0, v1, v2).tolist() == [1.1 + 3.3, 2.2 + 2.2, 3.3 + 1.0]
    assert O.vector_binop(1, v1, v2).tolist() == [1.1 - 3.3, 2.2 - 2.2, 3.3 - 1.0]


def test_vector_size_mismatch():  # test_vector.nim:41-45
    with pytest.raises(ValueError):
        O.vector_binop(0, [1.0, 2.0, 4.0, 1.34, 9.9], [3.3, 2.2, 1.1, 5.67])


def test_vector_scalar_ops():  # test_vector.nim:32-39, 90-109, 22-25, 283-287
    v, d = [1.1, 2.2, 3.3], 8.98
    assert O.vector_unop(4, d, v).tolist() == [1.1 + d, 2.2 + d, 3.3 + d]
    assert O.vector_unop(0, d, v).tolist() == [1.1 * d, 2.2 * d, 3.3 * d]
    assert O.vector_unop(1, d, v).tolist() == [1.1 / d, 2.2 / d, 3.3 / d]
    assert O.vector_unop(2, 0.0, [1.0, 2.5, -3.34]).tolist() == [-1.0, -2.5, 3.34]
    assert O.vector_unop(3, 0.0, [1.0, -2.5, -3.34]).tolist() == [1.0, 2.5, 3.34]


def test_vector_hadamard_and_dot():  # test_vector.nim:111-132
    v1, v2 = [1.0, 2.0, 3.0], [4.0, 5.0, 6.0]
    assert O.vector_binop(2, v1, v2).tolist() == [4.0, 10.0, 18.0]
    assert O.vector_binop(3, v1, v2).tolist() == [1.0 / 4.0, 2.0 / 5.0, 3.0 / 6.0]
    assert O.vector_dot(v1, v2) == 32.0


def test_vector_sum_is_signed_and_sequential():  # test_vector.nim:289-293; utils.nim:233-235
    assert O.vector_sum([1.0, -2.0, 3.0, -4.5]) == ((1.0 + -2.0) + 3.0) + -4.5
    big = [1e16, 1.0, -1e16, 1.0]
    assert O.vector_sum(big) == (((0.0 + 1e16) + 1.0) - 1e16) + 1.0


def test_vector_norms():  # test_vector.nim:162-166, 252-281
    v = [1.0, 2.0, 3.0, 4.0]
    assert O.vector_norm(v, 2) == math.sqrt(1.0 + 4.0 + 9.0 + 16.0)
    assert O.vector_norm(v, 1) == 10.0
    assert O.vector_norm(v, 0) == 4.0
    assert O.vector_norm(v, 4) == math.pow(1.0 + 16.0 + 81.0 + 256.0, 0.25)


# ---- tests/test_utils.nim ----------------------------------------------------------------------
def test_linspace_exact_integer_grid():  # test_utils.nim:15-23
    assert O.linspace(0.0, 10.0, 11).tolist() == [float(i) for i in range(11)]
    with pytest.raises(ValueError):
        O.linspace(0.0, 1.0, 0)


def test_is_close():  # test_utils.nim:5-13
    assert O.is_close(1.0, 1.0005, 1e-3)
    assert not O.is_close(1.0, 1.002, 1e-3)
    assert O.is_close([1.0, 2.0], [1.0, 2.0 + 1e-4], 1e-3)
