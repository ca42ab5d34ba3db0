"""Right-hand sides given as source (b200rk_jit_rhs_new): NVRTC compiles the caller's element-local expression into
the library's own kernels — the plain dy = f(t, y) kernel of the stage / RHS / finish pipeline, the whole-attempt
kernel, the device-resident driver loop and the one-kernel RK4 step. Compared with the CPU oracle driven by the same
expression as a numpy callback (numpy rounds every * / + separately, like the device code compiled with --fmad=false).

Bar as in test_gpu_parity.py: element-wise results bit-exact; quantities behind the error-norm reduction within
the stated tolerances; fixed-step trajectories bit-identical."""
import numpy as np
import pytest
from conftest import assert_bitwise_equal
from test_gpu_parity import ALL, RTOL_DT, RTOL_NORM, assert_states_close, norm_rtol

import oracle as O

pytestmark = pytest.mark.gpu

# (expression, host restatement) pairs; p = list of parameter arrays, c = list of scalars
LOGISTIC = ("c0*y*(1.0 - y/p0) + c1*t", lambda t, y, p, c: c[0] * y * (1.0 - y / p[0]) + c[1] * t)
FORCED = ("-(p0*y) + p1*(c0*t)", lambda t, y, p, c: -(p[0] * y) + p[1] * (c[0] * t))
SCALE = ("c0*y", lambda t, y, p, c: c[0] * y)


@pytest.fixture(scope="module")
def nn():
    import numericalnim_b200 as nn
    nn.default_context()
    return nn


@pytest.fixture(params=[1, 0], ids=["fused", "unfused"])
def fuse_mode(nn, request):
    ctx = nn.default_context()
    ctx.set("fuse_pointwise", request.param)
    yield request.param
    ctx.set("fuse_pointwise", 1)


def _problem(rng, n, case):
    y = 1.0 + 0.5 * rng.uniform(-1.0, 1.0, n)
    if case is LOGISTIC:
        return y, [rng.uniform(2.0, 5.0, n)], [0.7, 0.05]
    if case is FORCED:
        return y, [rng.uniform(0.1, 5.0, n), rng.uniform(-1.0, 1.0, n)], [0.3]
    return y, [], [-0.37]


def _both(nn, case, p, c):
    """(device right-hand side, oracle right-hand side) for the same expression."""
    expr, host = case
    vecs = [nn.newVector(a) for a in p]
    return nn.rhsJit(expr, vecs, c), O.rhs_callback(lambda t, y: host(t, y, p, c))


def _eval(rhs, t, gy):
    out = gy._new_like()
    assert rhs.fn(t, gy._h, out._h, rhs.user) == 0
    return out.to_numpy()


@pytest.mark.parametrize("l2", [0, 1], ids=["w2", "w4_l2"])
@pytest.mark.parametrize("case", [LOGISTIC, FORCED, SCALE], ids=["logistic", "forced", "scale"])
def test_jit_rhs_kernel_bitwise(nn, case, l2):
    """dy = f(t, y): both variants of the run-time compiled kernel (128-bit; 256-bit with the L2 hand-off hints)."""
    ctx = nn.default_context()
    rng = np.random.default_rng(5)
    try:
        ctx.set("l2_hints", l2)
        for n in [1, 2, 3, 4, 5, 1023, 2048, 65536 + 7]:
            y, p, c = _problem(rng, n, case)
            rhs, _ = _both(nn, case, p, c)
            got = _eval(rhs, 0.625, nn.newVector(y))
            assert_bitwise_equal(got, case[1](0.625, y, p, c), f"{case[0]} n={n}")
    finally:
        ctx.set("l2_hints", -1)


@pytest.mark.parametrize("case", [LOGISTIC, FORCED], ids=["logistic", "forced"])
@pytest.mark.parametrize("method", ALL)
def test_jit_single_step_matches_oracle(nn, method, case, fuse_mode):
    """One IntegratorProc call of every method with a t-dependent, nonlinear user expression."""
    rng = np.random.default_rng(41)
    n = 2049
    y, p, c = _problem(rng, n, case)
    rhs, orhs = _both(nn, case, p, c)
    t0, dt = 0.3, 0.005
    fsal = case[1](t0, y, p, c)
    opts = dict(absTol=1e-3, relTol=1e-3, dtMax=1.0, dtMin=1e-8)
    yn_ref, fn_ref, dt_ref, err_ref, st = O.step_vector(method, orhs, t0, y, fsal, dt, O.new_options(**opts))
    assert st.rejected == 0
    yn, fn, dt_used, err = nn.integratorStep(method, rhs, t0, nn.newVector(y), nn.newVector(fsal), dt, nn.newODEoptions(**opts))
    assert dt_used == dt_ref
    assert_bitwise_equal(yn.to_numpy(), yn_ref, f"{method} yNew")
    assert_bitwise_equal(fn.to_numpy(), fn_ref, f"{method} FSAL out")
    if err_ref == 0.0:
        assert err == 0.0
    else:
        assert abs(err - err_ref) <= RTOL_NORM * err_ref


@pytest.mark.parametrize("method", ["dopri54", "tsit54", "vern65"])
def test_jit_single_step_with_rejections(nn, method, fuse_mode):
    rng = np.random.default_rng(43)
    n = 1025
    y, p, c = _problem(rng, n, FORCED)
    p[0] = p[0] * 10.0
    rhs, orhs = _both(nn, FORCED, p, c)
    fsal = FORCED[1](0.0, y, p, c)
    opts = dict(absTol=1e-10, relTol=1e-10, dtMax=1.0, dtMin=1e-9)
    yn_ref, fn_ref, dt_ref, err_ref, st = O.step_vector(method, orhs, 0.0, y, fsal, 0.2, O.new_options(**opts))
    assert st.rejected >= 1
    yn, fn, dt_used, err = nn.integratorStep(method, rhs, 0.0, nn.newVector(y), nn.newVector(fsal), 0.2, nn.newODEoptions(**opts))
    assert abs(dt_used - dt_ref) <= RTOL_DT * dt_ref
    assert_states_close(yn.to_numpy(), yn_ref, method)
    assert abs(err - err_ref) <= 1e-8 * err_ref


@pytest.mark.parametrize("strict", [0, 1])
@pytest.mark.parametrize("vec_width", [4, 2])
@pytest.mark.parametrize("method", ["dopri54", "tsit54", "vern65", "rk4"])
def test_jit_fused_attempt_bitwise_equals_pipeline(nn, method, vec_width, strict):
    """The run-time compiled whole-attempt kernel (one launch) vs the same expression through the pipeline."""
    ctx = nn.default_context()
    rng = np.random.default_rng(79)
    out = {}
    try:
        ctx.set("strict_zeros", strict)
        ctx.set("vec_width", vec_width)
        for n in [1, 3, 4, 5, 1023, 65536 + 7]:
            y, p, c = _problem(rng, n, LOGISTIC)
            rhs, _ = _both(nn, LOGISTIC, p, c)
            gy = nn.newVector(y)
            fsal = nn.newVector(_eval(rhs, 0.25, gy))
            o = nn.newODEoptions(absTol=1e-4, relTol=1e-4, dtMax=1.0, dtMin=1e-8, dt=0.01)
            for fuse in (1, 0):
                ctx.set("fuse_pointwise", fuse)
                l0 = ctx.stats()["launches"]
                yn, fn, dt_used, err = nn.integratorStep(method, rhs, 0.25, gy, fsal, 0.01, o)
                out[fuse] = (yn.to_numpy(), fn.to_numpy(), dt_used, err, ctx.stats()["launches"] - l0)
            assert_bitwise_equal(out[1][0], out[0][0], f"{method} yNew n={n}")
            if method != "rk4":
                assert_bitwise_equal(out[1][1], out[0][1], f"{method} FSAL n={n}")
            assert out[1][2] == out[0][2]
            assert abs(out[1][3] - out[0][3]) <= norm_rtol(n) * abs(out[0][3])
            assert out[1][4] == 1 and out[0][4] >= 8  # one kernel instead of >= 8
    finally:
        ctx.set("strict_zeros", 0)
        ctx.set("vec_width", 4)
        ctx.set("fuse_pointwise", 1)


def test_jit_rk4_dense_backward_trajectory_bit_identical(nn, fuse_mode):
    """tspan on both sides of tStart with dense output: the backward pass evaluates g(t, y) = -f(-t, y)
    (ode.nim:545) with the user's t-dependent expression inside the fused RK4 kernel."""
    n = 513
    rng = np.random.default_rng(3)
    y0, p, c = _problem(rng, n, FORCED)
    rhs, orhs = _both(nn, FORCED, p, c)
    ts = nn.linspace(-1.0, 0.5, 6)
    t, ys = nn.solveODE(rhs, nn.newVector(y0), ts, nn.newODEoptions(dt=1e-2), integrator="rk4")
    ref = O.solve_vector("rk4", orhs, y0, ts, O.new_options(dt=1e-2))
    assert t == ref.t.tolist()
    assert_bitwise_equal(np.array([v.to_numpy() for v in ys]), ref.y, "rk4 + jit rhs, dense + backward")
    assert nn.ode.last_stats["rhs_evals"] == ref.stats.rhs_evals


@pytest.mark.parametrize("method", ["dopri54", "tsit54", "vern65", "bs32", "rk21"])
def test_jit_adaptive_dense_backward_matches_oracle(nn, method, fuse_mode):
    n = 257
    rng = np.random.default_rng(9)
    y0, p, c = _problem(rng, n, FORCED)
    rhs, orhs = _both(nn, FORCED, p, c)
    kw = dict(absTol=1e-7, relTol=1e-7, dtMax=0.25, dtMin=1e-8)
    ts = nn.linspace(-0.75, 1.0, 8)
    t, ys = nn.solveODE(rhs, nn.newVector(y0), ts, nn.newODEoptions(**kw), integrator=method)
    ref = O.solve_vector(method, orhs, y0, ts, O.new_options(**kw))
    st = nn.ode.last_stats
    assert t == ref.t.tolist()
    assert (st["steps"], st["rejected"], st["limiter_hits"], st["rhs_evals"]) == (ref.stats.steps, ref.stats.rejected, ref.stats.limiter_hits, ref.stats.rhs_evals)
    assert_states_close(np.array([v.to_numpy() for v in ys]), ref.y, method, rtol=1e-8 if method == "vern65" else 1e-9)


@pytest.mark.parametrize("case", [LOGISTIC, FORCED], ids=["logistic", "forced"])
@pytest.mark.parametrize("method", ["dopri54", "tsit54", "vern65"])
def test_jit_device_loop_step_sequence(nn, method, case):
    """The whole adaptive loop inside ONE persistent kernel built around the user's expression (stage times
    t + dt*c_s are formed on the device) vs the host-driven loop and the oracle."""
    ctx = nn.default_context()
    kw = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    rtol_dt = 1e-6 if method == "vern65" else 1e-10
    rng = np.random.default_rng(17)
    try:
        for n in [1, 5, 1000, 65536 + 3]:
            y0, p, c = _problem(rng, n, case)
            rhs, orhs = _both(nn, case, p, c)
            ref = O.solve_vector(method, orhs, y0, [0.0, 2.0], O.new_options(**kw), trace=True)
            seqs = {}
            for devloop in (1, 0):
                ctx.set("device_loop", devloop)
                s = nn.Solver(method, rhs, nn.newVector(y0), 2.0, nn.newODEoptions(**kw))
                ts = [0.0]
                l0 = ctx.stats()["launches"]
                while True:
                    done, fin = s.advance(3)
                    if done:
                        ts.append(s.state()[0])
                    if fin:
                        break
                seqs[devloop] = (np.array(ts), s.state()[3].to_numpy(), s.stats(), ctx.stats()["launches"] - l0)
                s.close()
            (t_dev, y_dev, st_dev, l_dev), (t_host, y_host, st_host, l_host) = seqs[1], seqs[0]
            for st in (st_dev, st_host):
                assert (st["steps"], st["rejected"], st["limiter_hits"]) == (ref.stats.steps, ref.stats.rejected, ref.stats.limiter_hits), (method, n, st)
            assert st_dev["rhs_evals"] == st_host["rhs_evals"] == ref.stats.rhs_evals
            assert l_dev == len(t_dev) - 1 and l_host == st_host["attempts"]  # one launch per advance() vs one per attempt
            t_ref = np.concatenate([[0.0], np.cumsum([r[1] for r in ref.trace])])[::3]
            if len(t_ref) < len(t_dev):
                t_ref = np.append(t_ref, 2.0)
            assert np.allclose(t_dev, t_ref, rtol=rtol_dt, atol=0), (method, n)
            assert np.allclose(t_dev, t_host, rtol=rtol_dt, atol=0), (method, n)
            assert_states_close(y_dev, ref.y[-1], f"{method} n={n} device loop vs oracle")
            assert_states_close(y_dev, y_host, f"{method} n={n} device loop vs host loop")
    finally:
        ctx.set("device_loop", -1)


def test_jit_diag_expression_equals_builtin_bitwise(nn):
    """`-(p0*y)` handed over as source is the built-in diag-linear right-hand side, bit for bit, on every path."""
    ctx = nn.default_context()
    n = 65536 + 5
    lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
    y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    kw = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    glam = nn.newVector(lam)
    try:
        for devloop in (0, 1):
            ctx.set("device_loop", devloop)
            res = []
            for rhs in (nn.rhsJit("-(p0*y)", [glam]), nn.rhsDiagLinear(glam)):
                t, ys = nn.solveODE(rhs, nn.newVector(y0), [0.0, 2.0], nn.newODEoptions(**kw), integrator="dopri54")
                res.append((ys[-1].to_numpy(), dict(nn.ode.last_stats)))
            assert_bitwise_equal(res[0][0], res[1][0], f"jit vs built-in, device_loop={devloop}")
            assert res[0][1] == res[1][1]
    finally:
        ctx.set("device_loop", -1)


def test_jit_set_scalars_without_recompiling(nn):
    n = 1000
    y = np.linspace(0.5, 1.5, n)
    rhs = nn.rhsJit("c0*y + c1", scalars=[2.0, 1.0])
    gy = nn.newVector(y)
    assert_bitwise_equal(_eval(rhs, 0.0, gy), 2.0 * y + 1.0)
    rhs.set_scalars([-3.0, 0.25])
    assert_bitwise_equal(_eval(rhs, 0.0, gy), -3.0 * y + 0.25)
    yn, _, _, _ = nn.integratorStep("rk4", rhs, 0.0, gy, None, 0.01, nn.newODEoptions(dt=0.01))
    ref, *_ = O.step_vector("rk4", O.rhs_callback(lambda t, v: -3.0 * v + 0.25), 0.0, y, y, 0.01, O.new_options(dt=0.01))
    assert_bitwise_equal(yn.to_numpy(), ref, "rk4 after set_scalars")
    with pytest.raises(ValueError, match="scalar count"):
        rhs.set_scalars([1.0])


def test_jit_errors(nn):
    with pytest.raises(ValueError, match=r'identifier "z" is undefined'):
        nn.rhsJit("z*y")
    a, b = nn.newVector(np.ones(8)), nn.newVector(np.ones(9))
    with pytest.raises(ValueError, match="same size"):
        nn.rhsJit("p0*p1*y", [a, b])
    rhs = nn.rhsJit("p0*y", [a])
    with pytest.raises(ValueError, match="same size"):  # utils.nim:22-26
        nn.integratorStep("dopri54", rhs, 0.0, b, b, 0.01)


def test_jit_math_functions_within_libm_tolerance(nn):
    """Transcendentals come from the CUDA math library, not glibc: agreement to a few ulps, not bit for bit."""
    n = 4097
    y = np.linspace(0.1, 2.0, n)
    rhs = nn.rhsJit("-y*exp(-c0*t) + sin(p0)", [nn.newVector(y[::-1].copy())], [0.5])
    got = _eval(rhs, 1.25, nn.newVector(y))
    ref = -y * np.exp(-0.5 * 1.25) + np.sin(y[::-1])
    assert np.max(np.abs(got - ref)) <= 4e-16 * np.max(np.abs(ref))
