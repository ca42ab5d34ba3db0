"""Property test: the C++ oracle and the independent pure-Python restatement agree BIT FOR BIT on randomly drawn
small IVPs — integrator, tolerances, dt bounds, tStart, and the shape of tspan (unsorted, duplicates, one-sided,
length 1 or 2) are all drawn. This widens the pinning of the oracle beyond the committed golden fixtures, in
particular on the rejection / dtMin-limiter branches no reference test reaches."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle as O
import pyref as R


@st.composite
def ivps(draw):
    n = draw(st.integers(1, 5))
    lam = [draw(st.floats(0.05, 60.0)) * draw(st.sampled_from([1.0, 1.0, 8.0])) for _ in range(n)]
    y0 = [draw(st.floats(0.1, 3.0)) * draw(st.sampled_from([1.0, -1.0])) for _ in range(n)]
    method = draw(st.sampled_from(["dopri54", "tsit54", "vern65", "rk4"]))
    tol = draw(st.sampled_from([1e-3, 1e-5, 1e-7, 1e-9, 1e-12]))
    dt_min = draw(st.sampled_from([1e-6, 1e-4, 1e-3, 4e-3]))  # the large ones force the dtMin limiter
    dt_max = draw(st.sampled_from([1e-2, 5e-2, 0.2]))
    t_start = draw(st.sampled_from([0.0, 0.25, -0.5]))
    k = draw(st.integers(1, 5))
    tspan = [draw(st.sampled_from([-0.6, -0.5, -0.3, -0.1, 0.0, 0.1, 0.25, 0.3, 0.45, 0.7])) for _ in range(k)]
    return method, lam, y0, dict(absTol=tol, relTol=tol, dtMax=dt_max, dtMin=dt_min, dt=dt_max / 4, tStart=t_start), tspan


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(ivps())
def test_oracle_equals_python_restatement_on_random_ivps(case):
    method, lam, y0, opts, tspan = case
    t_py, y_py, st_py = R.solve(R.rhs_diag(lam), R.Vec(y0), tspan, R.options(**opts), method)
    sol = O.solve_vector(method, O.rhs_diag_linear(lam), y0, tspan, O.new_options(**opts), trace=True)
    assert sol.t.tolist() == t_py
    y_or = sol.y
    assert y_or.shape[0] == len(y_py)
    for a, b in zip(y_or, y_py):
        assert np.array_equal(np.asarray(a).view(np.uint64), np.asarray(b.c, dtype=np.float64).view(np.uint64))
    assert sol.stats.steps == st_py["steps"]
    if method != "rk4":
        assert (sol.stats.attempts, sol.stats.rejected, sol.stats.limiter_hits) == (st_py["attempts"], st_py["rejected"], st_py["limiter_hits"])
        assert [r[1] for r in sol.trace] == [r[1] for r in st_py["trace"]]


def test_repeated_tstart_is_reported_once():
    """ode.nim:485-487: `if t0 in tspan` adds (tStart, y0) once however often tStart occurs in tspan, so both
    returned lists are shorter than tspan. (A case the property test above found in the oracle's front-end.)"""
    for tspan, want in (([0.0, 0.0], [0.0]), ([0.0, 0.0, 0.5], [0.0, 0.5]), ([0.0, -0.5, 0.0, 0.0], [-0.5, 0.0])):
        sol = O.solve_vector("dopri54", O.rhs_scale(-0.1), [1.0, 2.0], tspan, O.new_options())
        t_py, y_py, _ = R.solve(R.rhs_scale(-0.1), R.Vec([1.0, 2.0]), tspan, R.options(), "dopri54")
        assert sol.t.tolist() == t_py == want
        assert sol.y.shape[0] == len(y_py) == len(want)


def test_courtesy_fused_multithreaded_baseline_equals_the_port():
    """bench.py's courtesy CPU row (fused attempt, all host cores) is the same arithmetic per element as the port:
    identical step sequence counts, final state bit-identical when the error norms round alike, else within 1e-12."""
    import numpy as np
    import oracle as O
    n = 5000
    lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
    y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    kw = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    for method in ("dopri54", "tsit54", "vern65"):
        ref = O.solve_vector(method, O.rhs_diag_linear(lam), y0, [0.0, 2.0], O.new_options(**kw))
        for threads in (1, 3):
            y, st = O.fused_mt_solve_diag(method, lam, y0, 2.0, O.new_options(**kw), threads=threads)
            assert (st.steps, st.rejected, st.limiter_hits) == (ref.stats.steps, ref.stats.rejected, ref.stats.limiter_hits), (method, threads)
            assert np.allclose(y, ref.y[-1], rtol=1e-9, atol=1e-13)
