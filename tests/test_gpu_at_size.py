"""Parity at the sizes BASELINE.json's configurations are quoted on (run on the B200 box: pytest -m gpu).

  * config 3's shape: adaptive Tsit54 / DOPRI54 / Vern65 on the Lorenz-96 ring at N = 2^16 and 2^18, tspan = [0, 1], the
    config-3 initial condition, on the library's default path (the one-kernel attempt over overlapped
    tiles) and on the stage / RHS / finish pipeline every user closure gets — each against the CPU oracle's solveODE:
    step, attempt and rejection counts equal; states within the chaotic tolerance below.
  * config 2 whole: one complete DOPRI54 solve of the 2^23-dimensional diag-linear IVP over [0, 2] (the end-to-end
    workload of bench.py) against the oracle (~3 minutes of one CPU core): counts equal, every dt of the step
    sequence within RTOL_DT, final state within RTOL_Y / ATOL_Y.
  * config 4's shape on one GPU: Vern65 on the 2^21-dimensional diag-linear IVP, device-resident loop and host loop.

Tolerances. Element-wise arithmetic is bit-identical to the oracle for a given dt; what differs is the summation order
of the error norm (tree vs sequential), i.e. dt in its last bits. Lorenz-96 at F = 8 is chaotic (leading Lyapunov
exponent ~1.7 per time unit), so that difference is amplified by about e^1.7 over [0, 1]: states are compared at
RTOL_L96 = 1e-7 relative — five orders above the observed ~1e-12, three below the solver's own tolerance of 1e-6."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

OPTS = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
RTOL_L96 = 1e-7
RTOL_Y, ATOL_Y, RTOL_DT = 1e-9, 1e-13, 1e-10


@pytest.fixture(scope="module")
def nn():
    import numericalnim_b200 as nn
    nn.default_context()  # raises loudly if the CUDA library / device is missing
    return nn


def l96_y0(n):
    return 8.0 + 0.01 * np.sin(2.0 * np.pi * 37.0 * np.arange(n) / n)


_oracle_cache = {}


def oracle_l96(method, n):
    key = (method, n)
    if key not in _oracle_cache:
        _oracle_cache[key] = O.solve_vector(method, O.rhs_lorenz96(8.0), l96_y0(n), [0.0, 1.0], O.new_options(**OPTS))
    return _oracle_cache[key]


@pytest.mark.parametrize("path", ["default", "pipeline"])
@pytest.mark.parametrize("log2n", [16, 18])
@pytest.mark.parametrize("method", ["tsit54", "dopri54", "vern65"])
def test_adaptive_lorenz96_at_size(nn, method, log2n, path):
    ctx = nn.default_context()
    n = 1 << log2n
    ref = oracle_l96(method, n)
    fuse = 1 if path == "default" else 0
    try:
        ctx.set("fuse_stencil", fuse)
        ctx.set("fuse_stencil_attempt", fuse)
        l0 = ctx.stats()["launches"]
        t, ys = nn.solveODE(nn.rhsLorenz96(8.0), nn.newVector(l96_y0(n)), [0.0, 1.0], nn.newODEoptions(**OPTS), integrator=method)
        st = dict(nn.ode.last_stats)
        launches = ctx.stats()["launches"] - l0
    finally:
        ctx.set("fuse_stencil", 1)
        ctx.set("fuse_stencil_attempt", 1)
    assert t == [0.0, 1.0] and len(ys) == 2
    assert (st["steps"], st["attempts"], st["rejected"]) == (ref.stats.steps, ref.stats.attempts, ref.stats.rejected), (st, ref.stats.steps, ref.stats.rejected)
    got, exp = ys[-1].to_numpy(), np.asarray(ref.y[-1])
    rel = np.max(np.abs(got - exp) / np.abs(exp))
    assert rel <= RTOL_L96, (method, n, path, float(rel))
    if path == "default":
        assert launches <= st["attempts"] + 6, (launches, st["attempts"])   # one kernel per attempt (+ the start-up evaluations and copies)
    else:
        assert launches >= 7 * st["attempts"]


def test_whole_config2_solve_at_2p23_matches_oracle(nn):
    """The complete solveODE call bench.py's end-to-end leg times — DOPRI54, y' = -lambda .* y, N = 2^23, [0, 2] — on the
    default path (device-resident fused loop) AND on the host-driven fused loop, against the oracle port."""
    ctx = nn.default_context()
    n = 1 << 23
    i = np.arange(n, dtype=np.float64)
    lam = 0.1 + 9.9 * i / float(n - 1)
    y0 = 1.0 + 0.5 * np.sin(2.0 * np.pi * i / float(n))
    ref = O.solve_vector("dopri54", O.rhs_diag_linear(lam), y0, [0.0, 2.0], O.new_options(**OPTS), trace=True)
    exp = np.asarray(ref.y[-1])
    glam, gy0 = nn.newVector(lam), nn.newVector(y0)
    try:
        for devloop in (1, 0):
            ctx.set("device_loop", devloop)
            t, ys = nn.solveODE(nn.rhsDiagLinear(glam), gy0, [0.0, 2.0], nn.newODEoptions(**OPTS), integrator="dopri54")
            st = dict(nn.ode.last_stats)
            assert t == [0.0, 2.0]
            assert (st["steps"], st["attempts"], st["rejected"]) == (ref.stats.steps, ref.stats.attempts, ref.stats.rejected), (devloop, st)
            got = ys[-1].to_numpy()
            bound = RTOL_Y * np.abs(exp) + ATOL_Y * np.max(np.abs(exp))
            assert np.all(np.abs(got - exp) <= bound), (devloop, float(np.max(np.abs(got - exp) / np.abs(exp))))
            assert np.array_equal(ys[0].to_numpy().view(np.uint64), y0.view(np.uint64))   # the tStart slot is y0 itself (ode.nim:485-487)
        # the dt sequence, step by step, through the resumable solver (host-driven loop: one attempt per launch)
        ctx.set("device_loop", 0)
        s = nn.Solver("dopri54", nn.rhsDiagLinear(glam), gy0, 2.0, nn.newODEoptions(**OPTS))
        t_prev, dts = 0.0, []
        while True:
            done, fin = s.advance(1)
            t_now = s.state()[0]
            if done:
                dts.append(t_now - t_prev)
                t_prev = t_now
            if fin:
                break
        s.close()
        ref_dt = np.array([r[1] for r in ref.trace])   # (t, dt_used, error, attempts) per accepted step
        assert len(dts) == ref.stats.steps == len(ref_dt)
        # t_{k+1} - t_k recovers dt to ~1e-15 * t / dt relative; the last step is min(dt, tEnd - t) on both sides
        assert np.allclose(np.array(dts), ref_dt, rtol=1e-9, atol=0), (dts, list(ref_dt))
    finally:
        ctx.set("device_loop", -1)
        glam.free()
        gy0.free()


@pytest.mark.parametrize("devloop", [1, 0])
def test_vern65_diag_linear_2p21_matches_oracle(nn, devloop):
    ctx = nn.default_context()
    n = 1 << 21
    i = np.arange(n, dtype=np.float64)
    lam = 0.1 + 9.9 * i / float(n - 1)
    y0 = 1.0 + 0.5 * np.sin(2.0 * np.pi * i / float(n))
    if "vern65_diag" not in _oracle_cache:   # ~35 s of one CPU core: once for both loops
        _oracle_cache["vern65_diag"] = O.solve_vector("vern65", O.rhs_diag_linear(lam), y0, [0.0, 1.0], O.new_options(**OPTS))
    ref = _oracle_cache["vern65_diag"]
    exp = np.asarray(ref.y[-1])
    try:
        ctx.set("device_loop", devloop)
        t, ys = nn.solveODE(nn.rhsDiagLinear(nn.newVector(lam)), nn.newVector(y0), [0.0, 1.0], nn.newODEoptions(**OPTS), integrator="vern65")
        st = dict(nn.ode.last_stats)
    finally:
        ctx.set("device_loop", -1)
    assert (st["steps"], st["rejected"]) == (ref.stats.steps, ref.stats.rejected), st
    got = ys[-1].to_numpy()
    assert np.all(np.abs(got - exp) <= RTOL_Y * np.abs(exp) + ATOL_Y * np.max(np.abs(exp)))
