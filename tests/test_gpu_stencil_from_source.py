"""Stencil right-hand sides handed to the library as SOURCE (b200rk_jit_stencil_rhs_new): an ODEProc
closure whose dydt[i] reads a cyclic neighbourhood of y, compiled by NVRTC into a plain dydt = f(t, y) kernel and — for the
FSAL pairs — into the one-kernel attempt over overlapped tiles. The oracle is driven by the same expression as a numpy
callback (np.roll for the neighbours; Python floats are IEEE doubles without contraction, and the units are compiled with
--fmad=false), so element-wise results are compared BIT FOR BIT; adaptive solves at the tolerances of test_gpu_parity."""
import numpy as np
import pytest
from conftest import assert_bitwise_equal

import oracle as O

pytestmark = pytest.mark.gpu
KW = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)


@pytest.fixture(scope="module")
def nn():
    import numericalnim_b200 as nn
    nn.default_context()
    return nn


def Yr(y, d):
    return np.roll(y, -d)   # Y(d)[i] = y[(i + d) mod N]


CASES = {
    "lorenz96": dict(expr="((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", rl=2, rr=1, cs=[8.0], f=lambda t, y, p, c: ((Yr(y, 1) - Yr(y, -2)) * Yr(y, -1) - y) + c[0]),
    "diffusion": dict(expr="c0*((Y(-1) - 2.0*Y(0)) + Y(1))*p0 + c1*t", rl=1, rr=1, cs=[0.3, 0.05], np=1,
                      f=lambda t, y, p, c: c[0] * ((Yr(y, -1) - 2.0 * y) + Yr(y, 1)) * p[0] + c[1] * t),
    "upwind": dict(expr="-(c0*(Y(0) - Y(-1)))", rl=1, rr=0, cs=[0.7], f=lambda t, y, p, c: -(c[0] * (y - Yr(y, -1)))),
    "wide": dict(expr="((Y(-3) + Y(2)) - 2.0*Y(0))*c0 - Y(0)*Y(0)*Y(0)*c1", rl=3, rr=2, cs=[0.2, 0.01],
                 f=lambda t, y, p, c: ((Yr(y, -3) + Yr(y, 2)) - 2.0 * y) * c[0] - y * y * y * c[1]),
}
# extreme radii: right-only (HL = 0), the maximum 8 / 8 (overlap 48 + 48 of a 1024-wide tile for the 7-stage pairs), and 0 / 0 (element-local)
CASES.update({
    "right_only": dict(expr="c0*(Y(3) - Y(0))", rl=0, rr=3, cs=[0.4], f=lambda t, y, p, c: c[0] * (Yr(y, 3) - y)),
    "radius8": dict(expr="((Y(-8) + Y(8)) - 2.0*Y(0))*c0", rl=8, rr=8, cs=[0.05], f=lambda t, y, p, c: ((Yr(y, -8) + Yr(y, 8)) - 2.0 * y) * c[0]),
    "radius0": dict(expr="c0*Y(0) + c1*t", rl=0, rr=0, cs=[-0.3, 0.1], f=lambda t, y, p, c: c[0] * y + c[1] * t),
})
SIZES = [17, 19, 1003, 1004, 1005, 1023, 1024, 1025, 2009, 4099, 65536 + 3]


def make(nn, name, n, rng):
    cs = CASES[name]
    pvals = [1.0 + 0.5 * rng.uniform(-1, 1, n) for _ in range(cs.get("np", 0))]
    gp = [nn.newVector(p) for p in pvals]
    rhs = nn.rhsJitStencil(cs["expr"], cs["rl"], cs["rr"], gp, cs["cs"])
    orhs = O.rhs_callback(lambda t, y: cs["f"](t, np.asarray(y), pvals, cs["cs"]))
    return rhs, orhs, pvals


@pytest.mark.parametrize("name", sorted(CASES))
def test_stencil_rhs_kernel_bitwise(nn, name):
    """dydt = f(t, y) as the pipeline calls it: tiles, cyclic halos and the ragged tail at sizes around the 1024-element tile."""
    rng = np.random.default_rng(7)
    for n in SIZES:
        rhs, orhs, pvals = make(nn, name, n, rng)
        y = 8.0 + rng.uniform(-1, 1, n)
        gy = nn.newVector(y)
        out = gy._new_like()
        assert rhs.fn(0.37, gy._h, out._h, rhs.user) == 0
        exp = CASES[name]["f"](0.37, y, pvals, CASES[name]["cs"])
        assert_bitwise_equal(out.to_numpy(), exp, f"{name} n={n}")


@pytest.mark.parametrize("method", ["dopri54", "tsit54", "vern65"])
def test_lorenz96_from_source_equals_the_builtin_bit_for_bit(nn, method):
    """The Lorenz-96 expression given as source runs the NVRTC-compiled one-kernel attempt; yNew and the new FSAL must equal the
    built-in kernel's bits (and the oracle's), around every tile seam, fused (one kernel) and unfused (pipeline + stencil rhs kernel)."""
    ctx = nn.default_context()
    rng = np.random.default_rng(11)
    o = nn.newODEoptions(absTol=1e-2, relTol=1e-2, dtMax=1.0, dtMin=1e-8, dt=0.005)
    try:
        for n in SIZES:
            y = 8.0 + rng.uniform(-1, 1, n)
            fs = O.rhs_eval(O.rhs_lorenz96(8.0), 0.0, y)
            gy, gf = nn.newVector(y), nn.newVector(fs)
            src = nn.rhsJitStencil(CASES["lorenz96"]["expr"], 2, 1, [], [8.0])
            yb, fb, dtb, errb = nn.integratorStep(method, nn.rhsLorenz96(8.0), 0.0, gy, gf, 0.005, o)
            res = {}
            for fuse in (1, 0):
                ctx.set("fuse_stencil_attempt", fuse)
                l0 = ctx.stats()["launches"]
                yn, fn, dt_used, err = nn.integratorStep(method, src, 0.0, gy, gf, 0.005, o)
                res[fuse] = (yn.to_numpy(), fn.to_numpy(), dt_used, err, ctx.stats()["launches"] - l0)
            ctx.set("fuse_stencil_attempt", 1)
            for fuse in (1, 0):
                assert_bitwise_equal(res[fuse][0], yb.to_numpy(), f"{method} yNew n={n} fuse={fuse}")
                assert_bitwise_equal(res[fuse][1], fb.to_numpy(), f"{method} FSAL n={n} fuse={fuse}")
                assert res[fuse][2] == dtb
            assert res[1][4] == 1 and res[0][4] > 10, (res[1][4], res[0][4])
            yn_ref, fn_ref, *_ = O.step_vector(method, O.rhs_lorenz96(8.0), 0.0, y, fs, 0.005, O.new_options(absTol=1e-2, relTol=1e-2, dtMax=1.0, dtMin=1e-8, dt=0.005))
            assert_bitwise_equal(res[1][0], yn_ref, f"{method} yNew vs oracle n={n}")
    finally:
        ctx.set("fuse_stencil_attempt", 1)


@pytest.mark.parametrize("method", ["dopri54", "tsit54", "vern65"])
@pytest.mark.parametrize("name", ["diffusion", "upwind", "wide", "right_only", "radius8", "radius0"])
def test_one_step_of_a_user_stencil_is_bit_identical_to_the_oracle(nn, name, method):
    """Parameter vectors, explicit time dependence, one-sided and wide neighbourhoods: one IntegratorProc call, fused vs pipeline vs oracle."""
    ctx = nn.default_context()
    rng = np.random.default_rng(13)
    kw = dict(absTol=1e-2, relTol=1e-2, dtMax=1.0, dtMin=1e-8, dt=0.01)
    try:
        for n in (19, 1004, 1025, 4099):
            rhs, orhs, pvals = make(nn, name, n, rng)
            y = 1.0 + 0.5 * rng.uniform(-1, 1, n)
            fs = CASES[name]["f"](0.25, y, pvals, CASES[name]["cs"])
            gy, gf = nn.newVector(y), nn.newVector(fs)
            res = {}
            for fuse in (1, 0):
                ctx.set("fuse_stencil_attempt", fuse)
                yn, fn, dt_used, err = nn.integratorStep(method, rhs, 0.25, gy, gf, 0.01, nn.newODEoptions(**kw))
                res[fuse] = (yn.to_numpy(), fn.to_numpy(), dt_used, err)
            yn_ref, fn_ref, dt_ref, err_ref, st = O.step_vector(method, orhs, 0.25, y, fs, 0.01, O.new_options(**kw))
            assert st.rejected == 0
            for fuse in (1, 0):
                assert_bitwise_equal(res[fuse][0], yn_ref, f"{name} {method} yNew n={n} fuse={fuse}")
                assert_bitwise_equal(res[fuse][1], fn_ref, f"{name} {method} FSAL n={n} fuse={fuse}")
                assert res[fuse][2] == dt_ref and abs(res[fuse][3] - err_ref) <= 1e-12 * abs(err_ref)
    finally:
        ctx.set("fuse_stencil_attempt", 1)


@pytest.mark.parametrize("name,method", [("diffusion", "dopri54"), ("wide", "tsit54"), ("upwind", "vern65"), ("lorenz96", "tsit54")])
def test_adaptive_solve_with_dense_output_and_backward_time(nn, name, method):
    """solveODE over a tspan on both sides of tStart (backward pass g = -f(-t, y): NEG time and sign) with dense output: counts
    equal to the oracle's, states within the adaptive tolerance (chaotic tolerance for Lorenz-96)."""
    rng = np.random.default_rng(17)
    n = 5003
    rhs, orhs, pvals = make(nn, name, n, rng)
    y0 = (8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(n) / n)) if name == "lorenz96" else 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    ts = [-0.1, 0.0, 0.05, 0.2, 0.3]
    ref = O.solve_vector(method, orhs, y0, ts, O.new_options(**KW))
    t, ys = nn.solveODE(rhs, nn.newVector(y0), ts, nn.newODEoptions(**KW), integrator=method)
    st = dict(nn.ode.last_stats)
    assert list(t) == list(ref.t) and len(ys) == len(ref.y)
    assert st["steps"] == ref.stats.steps and st["rejected"] == ref.stats.rejected, (st, ref.stats.steps, ref.stats.rejected)
    rtol = 1e-7 if name == "lorenz96" else 1e-9
    for got, exp in zip(ys, ref.y):
        g, e = got.to_numpy(), np.asarray(exp)
        assert np.all(np.abs(g - e) <= rtol * np.abs(e) + 1e-13 * np.max(np.abs(e))), float(np.max(np.abs(g - e)))


@pytest.mark.parametrize("name", sorted(CASES))
def test_rk4_step_in_one_kernel_is_bit_identical(nn, name):
    """RK4 (ode.nim:180-189) of a stencil from source: k1..k4 and the final combine in ONE kernel over tiles that overlap by 4 * radius —
    bit-identical to the stage / RHS pipeline and to the oracle, around the tile seams, with dense output and a backward pass."""
    ctx = nn.default_context()
    rng = np.random.default_rng(23)
    try:
        for n in (19, 21, 1003, 1011, 1012, 1013, 1024, 2023, 2024, 4099, 65536 + 3):
            rhs, orhs, pvals = make(nn, name, n, rng)
            y = 8.0 + rng.uniform(-1, 1, n)
            gy = nn.newVector(y)
            res = {}
            for fuse in (1, 0):
                ctx.set("fuse_stencil_attempt", fuse)
                l0 = ctx.stats()["launches"]
                yn, fn, dt_used, err = nn.integratorStep("rk4", rhs, 0.4, gy, gy, 2e-3, nn.newODEoptions(dt=2e-3))
                res[fuse] = (yn.to_numpy(), ctx.stats()["launches"] - l0)
            assert_bitwise_equal(res[1][0], res[0][0], f"{name} rk4 n={n}")
            assert res[1][1] <= 2 and res[0][1] >= 8, (res[1][1], res[0][1])
            yn_ref, *_ = O.step_vector("rk4", orhs, 0.4, y, y, 2e-3, O.new_options(dt=2e-3))
            assert_bitwise_equal(res[1][0], yn_ref, f"{name} rk4 vs oracle n={n}")
        ctx.set("fuse_stencil_attempt", 1)
        n = 3001
        rhs, orhs, pvals = make(nn, name, n, rng)
        y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
        ts = nn.linspace(-0.02, 0.03, 7)
        t, ys = nn.solveODE(rhs, nn.newVector(y0), ts, nn.newODEoptions(dt=2e-3), integrator="rk4")
        ref = O.solve_vector("rk4", orhs, y0, ts, O.new_options(dt=2e-3))
        assert_bitwise_equal(np.array([v.to_numpy() for v in ys]), ref.y, f"{name} rk4 trajectory (dense output, backward pass)")
    finally:
        ctx.set("fuse_stencil_attempt", 1)


@pytest.mark.parametrize("method", ["rk4", "bs32", "heun2", "ssprk3"])
def test_methods_without_a_fused_form_run_the_pipeline_bit_identically(nn, method):
    rng = np.random.default_rng(19)
    n = 2051
    rhs, orhs, pvals = make(nn, "diffusion", n, rng)
    y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    kw = dict(dt=2e-3) if method != "bs32" else dict(KW)
    t, ys = nn.solveODE(rhs, nn.newVector(y0), [0.0, 0.05], nn.newODEoptions(**kw), integrator=method)
    ref = O.solve_vector(method, orhs, y0, [0.0, 0.05], O.new_options(**kw))
    if method == "bs32":
        assert np.allclose(ys[-1].to_numpy(), ref.y[-1], rtol=1e-9, atol=0)
    else:
        assert_bitwise_equal(ys[-1].to_numpy(), ref.y[-1], method)


def test_argument_errors(nn):
    with pytest.raises(ValueError):
        nn.rhsJitStencil("Y(0)", 9, 0)
    with pytest.raises(ValueError) as e:
        nn.rhsJitStencil("Y(-2) + Y(0)", 1, 1)
    assert "Y(-2)" in str(e.value)
    rhs = nn.rhsJitStencil("Y(-2) + Y(1)", 2, 1)
    with pytest.raises(ValueError):   # shorter than the stencil
        y = nn.newVector(np.ones(3))
        out = y._new_like()
        from numericalnim_b200 import _capi
        _capi.check(rhs.fn(0.0, y._h, out._h, rhs.user), nn.default_context().handle)
