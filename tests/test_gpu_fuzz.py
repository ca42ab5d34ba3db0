"""GPU property test: the product's solveODE against the CPU oracle on randomly drawn small IVPs (same strategy as
tests/test_oracle_fuzz.py, which pins the oracle on the independent Python restatement for the very same cases)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings

import oracle as O
from test_oracle_fuzz import ivps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nn():
    import numericalnim_b200 as m

    return m


def _check(nn, case):
    method, lam, y0, opts, tspan = case
    ref = O.solve_vector(method, O.rhs_diag_linear(lam), y0, tspan, O.new_options(**opts), trace=True)
    t, ys = nn.solveODE(nn.rhsDiagLinear(nn.newVector(lam)), nn.newVector(y0), tspan, nn.newODEoptions(**opts), integrator=method)
    st = dict(nn.ode.last_stats)
    assert t == ref.t.tolist(), case
    assert len(ys) == ref.y.shape[0], case
    assert st["steps"] == ref.stats.steps, case
    if method != "rk4":
        assert (st["attempts"], st["rejected"], st["limiter_hits"]) == (ref.stats.attempts, ref.stats.rejected, ref.stats.limiter_hits), case
    for a, b in zip(ys, ref.y):
        a = a.to_numpy()
        if method == "rk4":
            assert np.array_equal(a.view(np.uint64), np.asarray(b).view(np.uint64)), case  # fixed step: bit-identical
        else:
            # limiter-accepted steps (error > 1) on stiff components amplify the 1e-13 norm difference: rtol 1e-6
            scale = np.max(np.abs(b)) if b.size else 0.0
            assert np.all(np.abs(a - b) <= 1e-6 * np.abs(b) + 1e-12 * scale), (case, a, b)


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(ivps())
def test_solve_matches_oracle_on_random_ivps_fused(nn, case):
    _check(nn, case)


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(ivps())
def test_solve_matches_oracle_on_random_ivps_pipeline(nn, case):
    ctx = nn.default_context()
    old = (ctx.get("fuse_pointwise"), ctx.get("device_loop"))
    ctx.set("fuse_pointwise", 0)
    ctx.set("device_loop", 0)
    try:
        _check(nn, case)
    finally:
        ctx.set("fuse_pointwise", old[0])
        ctx.set("device_loop", old[1])


def test_repeated_tstart_is_reported_once(nn):
    """ode.nim:485-487: `if t0 in tspan` adds tStart once however often it occurs; the time list is then shorter
    than tspan. (Found by the property test above.)"""
    rhs = nn.rhsScale(-0.1)
    for tspan in ([0.0, 0.0], [0.0, 0.0, 0.5], [0.0, -0.5, 0.0, 0.0]):
        ref = O.solve_vector("dopri54", O.rhs_scale(-0.1), [1.0, 2.0], tspan, O.new_options())
        t, ys = nn.solveODE(rhs, nn.newVector([1.0, 2.0]), tspan, nn.newODEoptions(), integrator="dopri54")
        assert t == ref.t.tolist() and len(t) < len(tspan)
        assert len(ys) == ref.y.shape[0]
