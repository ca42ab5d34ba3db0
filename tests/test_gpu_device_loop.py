"""Device-resident driver loop (fused_run_kernel): many accepted steps of the adaptive loop — retry loop and
both controllers included — inside ONE persistent cooperative kernel. Compared with the host-driven loop and
with the CPU oracle. The device pow() may differ from glibc's in the last ulp, so step sequences are compared
within the tolerances of DESIGN.md §5; step / rejection / limiter counts must be equal."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nn():
    import numericalnim_b200 as nn
    nn.default_context()
    return nn


def _states_close(got, ref, rtol=1e-9):
    bound = rtol * np.abs(ref) + 1e-13 * np.max(np.abs(ref))
    return bool(np.all(np.abs(got - ref) <= bound))


@pytest.mark.parametrize("rhs_kind", ["diag", "scale"])
@pytest.mark.parametrize("method", ["dopri54", "tsit54", "vern65"])
def test_device_loop_step_sequence(nn, method, rhs_kind):
    ctx = nn.default_context()
    kw = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    rtol_dt = 1e-6 if method == "vern65" else 1e-10
    try:
        for n in [1, 5, 1000, 65536 + 3]:
            lam = 0.1 + 9.9 * np.arange(n) / max(n - 1, 1)
            y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
            orhs = O.rhs_diag_linear(lam) if rhs_kind == "diag" else O.rhs_scale(-1.7)
            ref = O.solve_vector(method, orhs, y0, [0.0, 2.0], O.new_options(**kw), trace=True)
            seqs = {}
            for devloop in (1, 0):
                ctx.set("device_loop", devloop)
                rhs = nn.rhsDiagLinear(nn.newVector(lam)) if rhs_kind == "diag" else nn.rhsScale(-1.7)
                s = nn.Solver(method, rhs, nn.newVector(y0), 2.0, nn.newODEoptions(**kw))
                ts = [0.0]
                l0 = ctx.stats()["launches"]
                while True:
                    done, fin = s.advance(1)
                    if done:
                        ts.append(s.state()[0])
                    if fin:
                        break
                seqs[devloop] = (np.array(ts), s.state()[3].to_numpy(), s.stats(), ctx.stats()["launches"] - l0)
                s.close()
            t_dev, y_dev, st_dev, _ = seqs[1]
            t_host, y_host, st_host, _ = seqs[0]
            for st in (st_dev, st_host):
                assert (st["steps"], st["rejected"], st["limiter_hits"]) == (ref.stats.steps, ref.stats.rejected, ref.stats.limiter_hits), (method, n, st)
            assert st_dev["rhs_evals"] == st_host["rhs_evals"]
            t_ref = np.concatenate([[0.0], np.cumsum([r[1] for r in ref.trace])])
            assert np.allclose(t_dev, t_ref, rtol=rtol_dt, atol=0), (method, n)
            assert np.allclose(t_dev, t_host, rtol=rtol_dt, atol=0), (method, n)
            assert _states_close(y_dev, ref.y[-1]) and _states_close(y_dev, y_host), (method, n)
    finally:
        ctx.set("device_loop", -1)


def test_whole_solve_is_one_kernel(nn):
    """tspan of length 2, element-local RHS, small N: solveODE runs its entire adaptive loop in one launch."""
    ctx = nn.default_context()
    n = 4096
    lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
    y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    kw = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    ref = O.solve_vector("dopri54", O.rhs_diag_linear(lam), y0, [0.0, 2.0], O.new_options(**kw))
    try:
        ctx.set("device_loop", 1)
        t, ys = nn.solveODE(nn.rhsDiagLinear(nn.newVector(lam)), nn.newVector(y0), [0.0, 2.0], nn.newODEoptions(**kw), integrator="dopri54")
        st = dict(nn.ode.last_stats)
        assert st["steps"] == ref.stats.steps and st["rejected"] == ref.stats.rejected
        assert st["launches"] <= 4, st  # 2 start-up RHS evaluations (ode.nim:498,506) + the loop kernel
        assert _states_close(ys[-1].to_numpy(), ref.y[-1])
    finally:
        ctx.set("device_loop", -1)


def test_device_loop_limiter_and_rejections(nn):
    """Stiff 4-vector with tight tolerance and a large dtMin: rejections and the dtMin limiter (ode.nim:71-76)
    run on the device; counts equal the oracle's."""
    ctx = nn.default_context()
    lam = np.array([1.0, 10.0, 100.0, 1000.0])
    kw = dict(absTol=1e-8, relTol=1e-8, dtMax=0.1, dtMin=1e-3)
    try:
        ctx.set("device_loop", 1)
        for method in ("dopri54", "tsit54", "vern65"):
            ref = O.solve_vector(method, O.rhs_diag_linear(lam), np.ones(4), [0.0, 0.05], O.new_options(**kw))
            t, ys = nn.solveODE(nn.rhsDiagLinear(nn.newVector(lam)), nn.newVector(np.ones(4)), [0.0, 0.05], nn.newODEoptions(**kw), integrator=method)
            st = dict(nn.ode.last_stats)
            assert (st["steps"], st["rejected"], st["limiter_hits"]) == (ref.stats.steps, ref.stats.rejected, ref.stats.limiter_hits), method
            assert ref.stats.limiter_hits > 0
            assert _states_close(ys[-1].to_numpy(), ref.y[-1], rtol=1e-6), method
    finally:
        ctx.set("device_loop", -1)


def test_device_loop_nan_is_reported(nn):
    from numericalnim_b200 import B200rkError
    ctx = nn.default_context()
    try:
        ctx.set("device_loop", 1)
        with pytest.raises(B200rkError) as ei:
            nn.solveODE(nn.rhsScale(-1.0), nn.newVector(np.array([1.0, np.nan, 2.0])), [0.0, 1.0], integrator="tsit54")
        assert ei.value.code == 6
    finally:
        ctx.set("device_loop", -1)
