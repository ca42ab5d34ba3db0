#!/usr/bin/env python
"""Tiny cases for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the library at n <= 4099, each case a
few launches, checked against the oracle so that a sanitizer pass also is a correctness pass. Usage:
    compute-sanitizer --tool memcheck python scripts/sanitizer_cases.py <case> [...]
Under torchrun (WORLD_SIZE > 1) the `sharded` case runs the in-kernel mailbox all-reduce and the peer-read stencil halo."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
KW = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)


def close(a, b, rtol=1e-9):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all(np.abs(a - b) <= rtol * np.abs(b) + 1e-13 * np.max(np.abs(b))))


def main():
    import numericalnim_b200 as nn
    import oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        from numericalnim_b200 import distributed as D
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
        ctx = D.init_context(int(os.environ["LOCAL_RANK"]))
    else:
        ctx = nn.default_context()
    nn.set_default_context(ctx)
    cases = sys.argv[1:] or ["pipeline", "fused", "device_loop", "l96", "quadrature", "jit"]
    ok = True

    def diag(n):
        return 0.1 + 9.9 * np.arange(n) / (n - 1), 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)

    def l96(n):
        return 8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(n) / n)

    def solve_diag(method, n, ts, **knobs):
        lam, y0 = diag(n)
        for k, v in knobs.items():
            ctx.set(k, v)
        g = nn.newVector(y0, ctx)
        t, ys = nn.solveODE(nn.rhsDiagLinear(nn.newVector(lam, ctx)), g, ts, nn.newODEoptions(**KW), integrator=method)
        ref = O.solve_vector(method, O.rhs_diag_linear(lam), y0, ts, O.new_options(**KW))
        st = dict(nn.ode.last_stats)
        lo, ln = g.local_offset, g.local_len
        return close(ys[-1].local_numpy(), ref.y[-1][lo:lo + ln]) and st["steps"] == ref.stats.steps, st

    for case in cases:
        if case == "pipeline":   # stage / RHS / finish kernels, element-wise operators, sum, dense output + backward pass
            r1, _ = solve_diag("dopri54", 1023, [0.0, 0.5], fuse_pointwise=0, device_loop=0)
            r2, _ = solve_diag("vern65", 1021, [-0.2, 0.0, 0.1, 0.3], fuse_pointwise=0, device_loop=0)
            r3, _ = solve_diag("rk4", 515, [0.0, 0.01], fuse_pointwise=0)
            a = nn.newVector(np.arange(1023.0), ctx)
            r4 = abs(((a + a) * 0.5 - a).sum()) == 0.0 and abs(a.sum() - 1023 * 1022 / 2) < 1e-6
            res = r1 and r2 and r3 and r4
            ctx.set("fuse_pointwise", 1)
            ctx.set("device_loop", -1)
        elif case == "fused":    # whole attempt in one kernel, host-driven loop
            res = all(solve_diag(m, 1023, [0.0, 0.5], device_loop=0)[0] for m in ("dopri54", "tsit54", "vern65", "rk4"))
            ctx.set("device_loop", -1)
        elif case == "device_loop":   # persistent cooperative kernel on SEVERAL CTAs: ticket + mailbox sum, no grid barrier
            r1, st1 = solve_diag("dopri54", 4099, [0.0, 0.5], device_loop=1)
            r2, st2 = solve_diag("vern65", 4099, [0.0, 0.5], device_loop=1)
            res = r1 and r2 and st1["launches"] <= 6
            ctx.set("device_loop", -1)
        elif case == "l96":      # one-kernel Lorenz-96 attempt (TMA bulk prefetch, shared-memory tiles), warp-tile variant, RK4 step, per-stage fusion
            res = True
            for knobs in (dict(), dict(l96_warp_tiles=1), dict(fuse_stencil_attempt=0), dict(fuse_stencil_attempt=0, fuse_stencil=0)):
                for k, v in knobs.items():
                    ctx.set(k, v)
                for method, ts, kw in (("tsit54", [0.0, 0.2], KW), ("vern65", [0.0, 0.2], KW), ("rk4", [0.0, 0.01], dict(dt=2e-3))):
                    y0 = l96(3001)
                    g = nn.newVector(y0, ctx)
                    t, ys = nn.solveODE(nn.rhsLorenz96(8.0, ctx), g, ts, nn.newODEoptions(**kw), integrator=method)
                    ref = O.solve_vector(method, O.rhs_lorenz96(8.0), y0, ts, O.new_options(**kw))
                    lo, ln = g.local_offset, g.local_len
                    res = res and close(ys[-1].local_numpy(), ref.y[-1][lo:lo + ln], 1e-7)
                for k, v in (("l96_warp_tiles", 0), ("fuse_stencil_attempt", 1), ("fuse_stencil", 1)):
                    ctx.set(k, v)
        elif case == "quadrature":
            rng = np.random.default_rng(3)
            X = np.sort(rng.uniform(0, 2, 9))
            Y = rng.uniform(-1, 1, (9, 1023))
            dv = [nn.newVector(r, ctx) for r in Y]
            lo, ln = dv[0].local_offset, dv[0].local_len
            res = True
            for name, ofn in (("cumtrapz", O.cumtrapz), ("cumsimpson", O.cumsimpson)):
                got = np.array([v.local_numpy() for v in getattr(nn, name)(dv, X)])
                res = res and np.array_equal(got.view(np.uint64), ofn(Y, X)[:, lo:lo + ln].view(np.uint64))
            xs = np.sort(rng.uniform(X[0], X[-1], 7))
            got = np.array([v.local_numpy() for v in nn.hermiteInterpolate(xs, X, dv, dv[::-1])])
            res = res and np.array_equal(got.view(np.uint64), O.hermite_interpolate(xs, X, Y, Y[::-1])[:, lo:lo + ln].view(np.uint64))
        elif case == "jit":      # right-hand side from source: NVRTC-compiled instances of the same kernels
            n = 1023
            K = 2.0 + 3.0 * np.arange(n) / (n - 1)
            y0 = diag(n)[1]
            rhs = nn.rhsJit("c0*y*(1.0 - y/p0) + c1*t", [nn.newVector(K, ctx)], [0.7, 0.05])
            ref = O.solve_vector("dopri54", O.rhs_callback(lambda t, y: 0.7 * y * (1.0 - y / K) + 0.05 * t), y0, [0.0, 1.0], O.new_options(**KW))
            res = True
            for dl in (1, 0):
                ctx.set("device_loop", dl)
                g = nn.newVector(y0, ctx)
                t, ys = nn.solveODE(rhs, g, [0.0, 1.0], nn.newODEoptions(**KW), integrator="dopri54")
                res = res and close(ys[-1].local_numpy(), ref.y[-1][g.local_offset:g.local_offset + g.local_len])
            ctx.set("device_loop", -1)
        elif case == "sharded":  # run under torchrun: mailbox all-reduce in both loops, stencil halo read from the peers
            r1, st1 = solve_diag("dopri54", 4099, [0.0, 0.5], device_loop=1)
            r2, st2 = solve_diag("tsit54", 4099, [0.0, 0.5], device_loop=0)
            y0 = l96(5003)
            g = nn.newVector(y0, ctx)
            t, ys = nn.solveODE(nn.rhsLorenz96(8.0, ctx), g, [0.0, 0.2], nn.newODEoptions(**KW), integrator="tsit54")
            ref = O.solve_vector("tsit54", O.rhs_lorenz96(8.0), y0, [0.0, 0.2], O.new_options(**KW))
            r3 = close(ys[-1].local_numpy(), ref.y[-1][g.local_offset:g.local_offset + g.local_len], 1e-7)
            res = r1 and r2 and r3 and (world == 1 or st1["collectives"] == st1["attempts"])
            ctx.set("device_loop", -1)
        else:
            raise SystemExit("unknown case " + case)
        print(f"[sanitizer case] {case}: {'ok' if res else 'WRONG RESULT'} (world={world}, rank={ctx.rank})", flush=True)
        ok = ok and res
    ctx.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
