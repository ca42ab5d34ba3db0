"""The fused paths that replaced multi-kernel pipelines in round 2 — the one-kernel Lorenz-96 attempt / RK4 step
(default since its first B200 run: 914 -> 3939 steps/s on config 3) and the single-pass cumsimpson (1.17 -> 0.67 ms) —
and the knobs that stayed off (finish_prefetch: no gain on the B200). Each test compares the fused kernel with the
multi-kernel pipeline it replaces (knob = 0) and with the oracle, bit for bit where the arithmetic is element-wise."""
import os

import numpy as np
import pytest

# AddressSanitizer pass of the host-emulated library (tests/test_host_emulated_library.py): the thread-per-CUDA-thread
# kernels are slow there, a few sizes around the tile seams are enough to catch an out-of-bounds access
LIGHT = bool(os.environ.get("B200RK_TEST_EMULATION_SANITIZE"))

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def nn():
    import numericalnim_b200 as nn
    nn.default_context()
    return nn


@pytest.mark.parametrize("vec_width", [4, 2])
def test_cumsimpson_single_pass_equals_two_kernel_path(nn, vec_width):
    """knob fuse_simpson=1: Simpson scan + Hermite interpolation in ONE kernel (simpson_fused_kernel), bit-identical to
    simpson_scan_kernel + hermite_many_kernel, one launch instead of two."""
    ctx = nn.default_context()
    rng = np.random.default_rng(5)
    try:
        ctx.set("vec_width", vec_width)
        for m in (3, 4, 5, 8, 17, 40):
            for n in (1, 3, 4, 5, 1023, 65536 + 7):
                X = np.sort(rng.uniform(0.0, 3.0, m))
                if m == 8:  # unsorted samples with a pure duplicate: results come back in the caller's order
                    X = np.concatenate([X[::-1], X[::-1][2:3]])
                Y = rng.uniform(-2.0, 2.0, (m, n))
                if m == 8:
                    Y = np.concatenate([Y[::-1], Y[::-1][2:3]])
                dv = [nn.newVector(r) for r in Y]
                res = {}
                for fuse in (0, 1):
                    ctx.set("fuse_simpson", fuse)
                    l0 = ctx.stats()["launches"]
                    out = nn.cumsimpson(dv, X)
                    res[fuse] = (np.array([v.to_numpy() for v in out]), ctx.stats()["launches"] - l0)
                assert res[0][0].shape == res[1][0].shape
                assert np.array_equal(res[0][0].view(np.uint64), res[1][0].view(np.uint64)), (m, n)
                if m != 8:
                    assert res[1][1] == 1 and res[0][1] == 2
    finally:
        ctx.set("fuse_simpson", 1)
        ctx.set("vec_width", 4)


@pytest.mark.parametrize("vec_width", [4, 2])
@pytest.mark.parametrize("method,stages", [("dopri54", 7), ("tsit54", 7), ("vern65", 9)])
def test_prefetching_finish_kernel_equals_default(nn, method, stages, vec_width):
    """knob finish_prefetch=1: the software-pipelined finish kernel (finish_pf.cuh) — same yNew / error_y bits and, on the
    same grid, the same error-norm bits as finish_kernel; whole solves take the identical step sequence."""
    ctx = nn.default_context()
    rng = np.random.default_rng(13)
    try:
        ctx.set("vec_width", vec_width)
        for n in (1, 3, 4, 5, 1023, 65536 + 7, (1 << 20) + 1):
            y = 1.0 + 0.5 * rng.uniform(-1.0, 1.0, n)
            ks = [nn.newVector(rng.uniform(-1.0, 1.0, n)) for _ in range(stages)]
            gy = nn.newVector(y)
            res = {}
            for pf in (0, 1):
                ctx.set("finish_prefetch", pf)
                yn, ey, S, E = nn.combineErr(method, 0.01, 1e-6, 1e-6, gy, ks, want_err_y=True)
                res[pf] = (yn.to_numpy(), ey.to_numpy(), S, E)
            assert np.array_equal(res[0][0].view(np.uint64), res[1][0].view(np.uint64)), (method, n)
            assert np.array_equal(res[0][1].view(np.uint64), res[1][1].view(np.uint64)), (method, n)
            assert res[0][2] == res[1][2] and res[0][3] == res[1][3], (method, n, res[0][2], res[1][2])
        # a whole adaptive solve on the general pipeline (fused paths off) takes the same steps either way
        n = 4096
        lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
        y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
        ctx.set("fuse_pointwise", 0)
        out = {}
        for pf in (0, 1):
            ctx.set("finish_prefetch", pf)
            t, ys = nn.solveODE(nn.rhsDiagLinear(nn.newVector(lam)), nn.newVector(y0), [0.0, 2.0],
                                nn.newODEoptions(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8), integrator=method)
            out[pf] = (ys[-1].to_numpy(), dict(nn.ode.last_stats))
        assert np.array_equal(out[0][0].view(np.uint64), out[1][0].view(np.uint64))
        assert out[0][1]["steps"] == out[1][1]["steps"] and out[0][1]["launches"] == out[1][1]["launches"]
    finally:
        ctx.set("finish_prefetch", 0)
        ctx.set("fuse_pointwise", 1)
        ctx.set("vec_width", 4)


@pytest.mark.parametrize("dx", [0.1, 0.013, 0.5])
def test_streaming_cumsimpson_fn_equals_composed_form(nn, dx):
    """knob stream_simpson=1: cumsimpson(f, X, dx) walks the grid through a window of a few vectors instead of keeping
    every evaluation alive — same results bit for bit, same callbacks in the same order. (Without the knob it is used
    only where the composed form would not fit or the grid has more than 4096 points.)"""
    import math

    import oracle as O
    ctx = nn.default_context()
    n = 4099
    a = np.linspace(1.0, 3.0, n)
    ga = nn.newVector(a)
    try:
        for X in (O.linspace(0.0, 1.5 * math.pi, 17), O.linspace(0.0, 1.5 * math.pi, 17)[[3, 0, 16, 7, 7, 1]], np.array([1.0, 2.0, 1.5])):
            ref, evals = O.cumsimpson_fn(lambda x: a * math.cos(x) + 0.1 * x, X, dx=dx, n=n)
            for mode in (0, 1):
                ctx.set("stream_simpson", mode)
                calls = []

                def f(x, c):
                    calls.append(x)
                    return math.cos(x) * ga + 0.1 * x

                got = np.array([v.to_numpy() for v in nn.cumsimpson(f, X, dx=dx, like=ga)])
                assert got.shape == ref.shape and np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (dx, mode)
                assert calls == sorted(calls) and len(calls) == evals
    finally:
        ctx.set("stream_simpson", -1)


def test_cumsimpson_fn_streams_a_fine_grid_by_itself(nn):
    """23,564 grid points (dx = 2e-4): beyond the composed form's limit, so the library streams on its own."""
    import math

    import oracle as O
    n = 64
    a = np.linspace(1.0, 3.0, n)
    ga = nn.newVector(a)
    X = O.linspace(0.0, 1.5 * math.pi, 17)
    got = np.array([v.to_numpy() for v in nn.cumsimpson(lambda x, c: math.cos(x) * ga, X, dx=2e-4, like=ga)])
    ref, _ = O.cumsimpson_fn(lambda x: a * math.cos(x), X, dx=2e-4, n=n)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    assert np.max(np.abs(got - np.outer(np.sin(X), a))) < 1e-10


# ---- knob fuse_stencil_attempt: a whole attempt of the built-in Lorenz-96 right-hand side in one kernel ---------------
@pytest.mark.parametrize("pairs", [2, 1, 0, -4, -128])   # -128: the CTA-tile kernel with 128-thread CTAs (knob l96_attempt_threads)   # 0 / -4: the warp-tile variant (knob l96_warp_tiles: shuffles, no shared-memory exchange; 8 / 4 elements per lane = 256- / 128-position tiles)
@pytest.mark.parametrize("method,stages", [("dopri54", 7), ("tsit54", 7), ("vern65", 9)])
def test_l96_attempt_kernel_bitwise_equals_pipeline(nn, method, stages, pairs):
    """l96_attempt_kernel (stencil_attempt.cuh: overlapped tiles, stage inputs through shared memory): yNew and the new
    FSAL carry the same bits as the stage_l96_kernel / finish_kernel pipeline and as the oracle, at sizes around the
    tile seams (1004 / 1000 stored elements per 1024-element tile; 492 / 488 per 512-element tile with l96_attempt_pairs=1) and below one tile (the tile wraps the cyclic domain
    several times), with and without the zero weights; one launch instead of S. (Backward time: the solve test below.)"""
    import os

    import oracle as O
    ctx = nn.default_context()
    rng = np.random.default_rng(17)
    out_per_tile = (512 * pairs if pairs > 0 else (256 if pairs == 0 else (512 if pairs == -128 else 128))) - (12 + 8 if stages == 7 else 16 + 8)
    sizes = [4, 5, 7, 19, 21, out_per_tile - 1, out_per_tile, out_per_tile + 1, 1023, 1024, 1025, 2 * out_per_tile - 1, 2 * out_per_tile,
             2 * out_per_tile + 2, 3 * out_per_tile + 13]
    if not os.environ.get("B200RK_TEST_HOST_EMULATION"):
        sizes += [8192 + 5, (1 << 20) + 7]   # on the GPU also sizes with thousands of tiles
    if LIGHT or (os.environ.get("B200RK_TEST_HOST_EMULATION") and pairs <= 0):   # the variant geometries: a few sizes around the seams are enough under the (slow) thread-per-thread emulation
        sizes = [5, out_per_tile - 1, out_per_tile, out_per_tile + 1, 1025, 2 * out_per_tile + 2]
    opts = dict(absTol=1e-2, relTol=1e-2, dtMax=1.0, dtMin=1e-8, dt=0.005)
    o = nn.newODEoptions(**opts)
    rhs = nn.rhsLorenz96(8.0)
    try:
        ctx.set("l96_attempt_pairs", pairs if pairs > 0 else 2)
        ctx.set("l96_warp_tiles", 8 if pairs == 0 else (4 if pairs == -4 else 0))
        ctx.set("l96_attempt_threads", 128 if pairs == -128 else 256)   # (128 is the default; the 256-thread geometry is what pairs = 2 / 1 test)
        for strict in (0, 1):
            ctx.set("strict_zeros", strict)
            for n in (sizes if not strict else sizes[5:9] if len(sizes) > 9 else sizes[2:3]):
                y = 8.0 + rng.uniform(-1.0, 1.0, n)
                fs = O.rhs_eval(O.rhs_lorenz96(8.0), 0.0, y)
                gy, gf = nn.newVector(y), nn.newVector(fs)
                res = {}
                for fuse in (1, 0):
                    ctx.set("fuse_stencil_attempt", fuse)
                    l0 = ctx.stats()["launches"]
                    yn, fn, dt_used, err = nn.integratorStep(method, rhs, 0.0, gy, gf, 0.005, o)
                    res[fuse] = (yn.to_numpy(), fn.to_numpy(), dt_used, err, ctx.stats()["launches"] - l0)
                assert np.array_equal(res[1][0].view(np.uint64), res[0][0].view(np.uint64)), (method, n, "yNew")
                assert np.array_equal(res[1][1].view(np.uint64), res[0][1].view(np.uint64)), (method, n, "FSAL")
                assert res[1][2] == res[0][2] and abs(res[1][3] - res[0][3]) <= 1e-13 * res[0][3]   # same terms, another order of summation
                assert res[1][4] == 1 and res[0][4] == stages
                if not strict:
                    yn_ref, fn_ref, _, err_ref, st = O.step_vector(method, O.rhs_lorenz96(8.0), 0.0, y, fs, 0.005, O.new_options(**opts))
                    assert st.rejected == 0
                    assert np.array_equal(res[1][0].view(np.uint64), yn_ref.view(np.uint64)), (method, n, "yNew vs oracle")
                    assert np.array_equal(res[1][1].view(np.uint64), fn_ref.view(np.uint64)), (method, n, "FSAL vs oracle")
                    assert abs(res[1][3] - err_ref) <= 1e-12 * err_ref
    finally:
        ctx.set("strict_zeros", 0)
        ctx.set("fuse_stencil_attempt", 1)
        ctx.set("l96_attempt_pairs", 2)
        ctx.set("l96_warp_tiles", 0)
        ctx.set("l96_attempt_threads", 128)


@pytest.mark.parametrize("method", ["tsit54", "vern65"])
def test_l96_attempt_solve_matches_pipeline_and_oracle(nn, method):
    """solveODE over a tspan on both sides of tStart with dense output, tolerances that make the controller reject:
    same accepted / rejected counts as the pipeline and the oracle, states within 1e-9 of the state's scale
    (the error norms differ in the last bits only, so the step sizes do too, and the chaotic system amplifies that: the
    default pipeline sits at the same 5e-11 from the oracle)."""
    import oracle as O
    ctx = nn.default_context()
    n = 2500 if not LIGHT else 1100
    y0 = 8.0 + np.random.default_rng(3).uniform(-4.0, 4.0, n)   # rough state: the controller overshoots and rejects
    ts = nn.linspace(-0.1, 0.3, 5) if not LIGHT else nn.linspace(-0.05, 0.1, 4)
    opts = dict(absTol=1e-5, relTol=1e-5, dtMax=1.0, dtMin=1e-4, tStart=0.0)
    res = {}
    try:
        for fuse in (1, 0):
            ctx.set("fuse_stencil_attempt", fuse)
            l0 = ctx.stats()["launches"]
            t, ys = nn.solveODE(nn.rhsLorenz96(8.0), nn.newVector(y0), ts, nn.newODEoptions(**opts), integrator=method)
            res[fuse] = (list(t), np.array([v.to_numpy() for v in ys]), dict(nn.ode.last_stats), ctx.stats()["launches"] - l0)
        ref = O.solve_vector(method, O.rhs_lorenz96(8.0), y0, ts, O.new_options(**opts))
        for fuse in (1, 0):
            assert res[fuse][0] == list(ref.t)
            assert res[fuse][2]["steps"] == ref.stats.steps and res[fuse][2]["rejected"] == ref.stats.rejected, (fuse, res[fuse][2])
            assert np.max(np.abs(res[fuse][1] - ref.y)) <= 1e-9 * np.max(np.abs(ref.y)), fuse
        assert LIGHT or ref.stats.rejected > 0
        assert np.max(np.abs(res[1][1] - res[0][1])) <= 1e-9 * np.max(np.abs(res[0][1]))
        assert res[1][3] < res[0][3] / 3
    finally:
        ctx.set("fuse_stencil_attempt", 1)


def test_l96_attempt_matches_golden_lorenz96_fixtures(nn):
    """The committed Lorenz-96 trajectories (tests/golden/trajectories.json: N = 40, far smaller than a tile, so every tile
    position wraps the ring ~25 times; several rejected attempts) with the one-kernel attempt switched on: same times,
    same step / attempt / rejection counts, states within the tolerance the default path is held to."""
    import json
    import os
    ctx = nn.default_context()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trajectories.json")) as fh:
        golden = json.load(fh)
    unhex = lambda xs: np.array([float.fromhex(x) for x in xs], dtype=np.float64)
    checked = 0
    try:
        ctx.set("fuse_stencil_attempt", 1)
        for name, g in golden.items():
            if g["rhs"]["kind"] != "l96" or g["integrator"] not in ("dopri54", "tsit54", "vern65"):
                continue
            l0 = ctx.stats()["launches"]
            t, ys = nn.solveODE(nn.rhsLorenz96(g["rhs"]["F"]), nn.newVector(unhex(g["y0"])), unhex(g["tspan"]), nn.newODEoptions(**g["options"]),
                                integrator=g["integrator"])
            st = dict(nn.ode.last_stats)
            assert np.array_equal(np.array(t).view(np.uint64), unhex(g["t"]).view(np.uint64)), name
            assert (st["steps"], st["attempts"], st["rejected"], st["limiter_hits"]) == (g["steps"], g["attempts"], g["rejected"], g["limiter_hits"]), (name, st)
            got, want = np.array([v.to_numpy() for v in ys]), np.array([unhex(r) for r in g["y"]])
            assert got.shape == want.shape and np.all(np.abs(got - want) <= 1e-6 * np.abs(want) + 1e-13 * np.max(np.abs(want))), name
            assert ctx.stats()["launches"] - l0 < 3 * g["attempts"] + 16   # one kernel per attempt (+ dense output / initial evaluations)
            checked += 1
        assert checked >= 1
    finally:
        ctx.set("fuse_stencil_attempt", 1)


def test_l96_rk4_step_in_one_kernel_is_bit_identical(nn):
    """knob fuse_stencil_attempt with RK4 (l96_rk4_kernel: k1..k4 and the final combine of one step over overlapped
    tiles, 8 + 4 elements of overlap): fixed step, so whole trajectories — dense output on both sides of tStart included —
    carry the same bits as the stage+stencil pipeline and the oracle; one launch per step instead of five."""
    import oracle as O
    ctx = nn.default_context()
    rng = np.random.default_rng(23)
    o = nn.newODEoptions(dt=0.005)
    rhs = nn.rhsLorenz96(8.0)
    try:
        for n in ([4, 5, 7, 21, 1011, 1012, 1013, 1024, 2024, 2025, 3049] if not LIGHT else [5, 1012, 1013, 2025]):
            y = 8.0 + rng.uniform(-1.0, 1.0, n)
            gy = nn.newVector(y)
            res = {}
            for fuse in (1, 0):
                ctx.set("fuse_stencil_attempt", fuse)
                l0 = ctx.stats()["launches"]
                yn, fn, dt_used, err = nn.integratorStep("rk4", rhs, 0.0, gy, None, 0.005, o)
                res[fuse] = (yn.to_numpy(), fn.to_numpy(), ctx.stats()["launches"] - l0)
            ref = O.step_vector("rk4", O.rhs_lorenz96(8.0), 0.0, y, y, 0.005, O.new_options(dt=0.005))[0]
            for a, b, what in ((res[1][0], res[0][0], "yNew vs pipeline"), (res[1][1], res[0][1], "returned FSAL vs pipeline"), (res[1][0], ref, "yNew vs oracle")):
                assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (n, what)
            assert res[1][2] < res[0][2]
        n = 1500
        y0 = 8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(n) / n)
        ts = nn.linspace(-0.05, 0.1, 7) if not LIGHT else nn.linspace(-0.01, 0.02, 4)
        out = {}
        for fuse in (1, 0):
            ctx.set("fuse_stencil_attempt", fuse)
            t, ys = nn.solveODE(rhs, nn.newVector(y0), ts, nn.newODEoptions(dt=2e-3), integrator="rk4")
            out[fuse] = np.array([v.to_numpy() for v in ys])
        want = O.solve_vector("rk4", O.rhs_lorenz96(8.0), y0, ts, O.new_options(dt=2e-3)).y
        assert np.array_equal(out[1].view(np.uint64), out[0].view(np.uint64))
        assert np.array_equal(out[1].view(np.uint64), np.asarray(want).view(np.uint64))
    finally:
        ctx.set("fuse_stencil_attempt", 1)


def test_round2_knobs_and_host_side_tstart_copy(nn):
    """Knobs added in round 2 are readable / writable, and solve_host returns the same bits whichever way the tStart slot is filled
    (D2H on a second stream, or a host-side copy of y0 by helper threads)."""
    import ctypes as C

    from numericalnim_b200 import _capi
    ctx = nn.default_context()
    for key, val in (("peer_timeout_s", 30), ("l96_ctas_per_sm", 2), ("l96_warp_tiles", 4), ("tstart_copy", 1)):
        old = ctx.get(key)
        ctx.set(key, val)
        assert ctx.get(key) == val
        ctx.set(key, old)
    with pytest.raises(ValueError):
        ctx.set("peer_timeout_s", 0)
    n = 70001
    lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
    y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    rhs = nn.rhsDiagLinear(nn.newVector(lam))
    ts = np.array([0.0, 0.5])
    opts = nn.newODEoptions(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    outs = {}
    try:
        for mode in (0, 1):
            ctx.set("tstart_copy", mode)
            out = np.empty((2, n))
            t_out = np.empty(2)
            n_out = C.c_size_t(0)
            st = _capi.Stats()
            _capi.check(_capi.lib().b200rk_solve_host(ctx.handle, nn.ode.method_id("dopri54"), rhs.fn, rhs.user, n, y0.ctypes.data, ts.ctypes.data, 2,
                                                      C.byref(opts), t_out.ctypes.data, out.ctypes.data, C.byref(n_out), C.byref(st)), ctx.handle)
            assert n_out.value == 2 and list(t_out) == [0.0, 0.5]
            outs[mode] = out
        assert np.array_equal(outs[0].view(np.uint64), outs[1].view(np.uint64))
        assert np.array_equal(outs[0][0].view(np.uint64), y0.view(np.uint64))
    finally:
        ctx.set("tstart_copy", 0)
