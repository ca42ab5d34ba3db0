#!/usr/bin/env python
"""Sharded-path check, one process per GPU (torchrun): the state is split contiguously across ranks, the
only collective is the library's own ncclAllReduce of the squared error norm. Compares every rank's shard
with the unsharded CPU oracle: same step count, states within the adaptive tolerance; RK4 bit-identical."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    import numericalnim_b200 as nn
    import oracle as O
    from numericalnim_b200 import distributed as D

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = D.init_context(local)
    assert (ctx.rank, ctx.world) == (rank, world)
    if rank == 0:
        print(f"[multi-gpu world={world}] error-norm all-reduce: {'in-kernel peer mailboxes (NVLink, CUDA IPC)' if ctx.get('p2p') else 'ncclAllReduce'}", flush=True)
    n = 100003
    lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
    y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    kw = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    glam, gy0 = nn.newVector(lam, ctx), nn.newVector(y0, ctx)
    off, ln = gy0.local_offset, gy0.local_len
    assert (off, ln) == D.shard_range(n, rank, world)
    fails = []
    for method in ("dopri54", "tsit54", "vern65", "rk4"):
        okw = dict(kw, dt=5e-3)
        ts = [0.0, 2.0] if method != "rk4" else [0.0, 0.5]
        ref = O.solve_vector(method, O.rhs_diag_linear(lam), y0, ts, O.new_options(**okw))
        t, ys = nn.solveODE(nn.rhsDiagLinear(glam), gy0, ts, nn.newODEoptions(**okw), integrator=method)
        st = dict(nn.ode.last_stats)
        got = ys[-1].local_numpy()
        exp = ref.y[-1][off:off + ln]
        if method == "rk4":
            ok = np.array_equal(got.view(np.uint64), exp.view(np.uint64)) and st["collectives"] == 0
        else:
            ok = bool(np.all(np.abs(got - exp) <= 1e-9 * np.abs(exp) + 1e-13 * np.max(np.abs(ref.y[-1])))) and st["collectives"] == st["attempts"]
        ok = ok and st["steps"] == ref.stats.steps and st["rejected"] == ref.stats.rejected
        if not ok:
            fails.append((method, st, ref.stats.steps, float(np.max(np.abs(got - exp)))))
        if rank == 0:
            print(f"[multi-gpu world={world}] {method}: steps={st['steps']} attempts={st['attempts']} collectives={st['collectives']} ok={ok}", flush=True)
    # the fused element-local path and the pipeline must agree bit for bit on every shard (fixed step)
    res = {}
    for fuse in (1, 0):
        ctx.set("fuse_pointwise", fuse)
        t, ys = nn.solveODE(nn.rhsDiagLinear(glam), gy0, [0.0, 0.2], nn.newODEoptions(dt=5e-3), integrator="rk4")
        res[fuse] = ys[-1].local_numpy()
    ctx.set("fuse_pointwise", 1)
    if not np.array_equal(res[0].view(np.uint64), res[1].view(np.uint64)):
        fails.append(("rk4 fused vs pipeline",))
    # sharded Lorenz-96: 3-element ring halo per right-hand-side evaluation (ncclSend/ncclRecv on the stream)
    for nl in (1000, 100003):
        yl = 8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(nl) / nl)
        gl = nn.newVector(yl, ctx)
        lo, ll = gl.local_offset, gl.local_len
        rhs_l = nn.rhsLorenz96(8.0, ctx)
        out = gl._new_like()
        assert rhs_l.fn(0.0, gl._h, out._h, rhs_l.user) == 0
        exp_k = O.rhs_eval(O.rhs_lorenz96(8.0), 0.0, yl)[lo:lo + ll]
        if not np.array_equal(out.local_numpy().view(np.uint64), exp_k.view(np.uint64)):
            fails.append(("lorenz96 rhs bitwise", nl))
        refl = O.solve_vector("tsit54", O.rhs_lorenz96(8.0), yl, [0.0, 0.5], O.new_options(**kw))
        t, ys = nn.solveODE(rhs_l, gl, [0.0, 0.5], nn.newODEoptions(**kw), integrator="tsit54")
        st = dict(nn.ode.last_stats)
        got, exp = ys[-1].local_numpy(), refl.y[-1][lo:lo + ll]
        ok = bool(np.all(np.abs(got - exp) <= 1e-7 * np.abs(exp))) and st["steps"] == refl.stats.steps and st["rejected"] == refl.stats.rejected
        if not ok:
            fails.append(("lorenz96 tsit54", nl, st, refl.stats.steps, float(np.max(np.abs(got - exp)))))
        if rank == 0:
            print(f"[multi-gpu world={world}] lorenz96 n={nl} tsit54: steps={st['steps']} rejected={st['rejected']} collectives={st['collectives']} ok={ok}", flush=True)
        # experimental knob fuse_stencil_attempt: the whole attempt in one kernel; sharded, ONE halo exchange of y and k1
        # (12 + 8 elements each) per IntegratorProc call instead of a 3-element exchange per right-hand-side evaluation
        # halo: read in place from the peer-mapped neighbours (needs the mailboxes' IPC path) or one ncclSend/ncclRecv per step
        try:
            ctx.set("fuse_stencil_attempt", 1)
            for peer_halo in (1, 0):
                ctx.set("l96_peer_halo", peer_halo)
                t, ys = nn.solveODE(rhs_l, gl, [0.0, 0.5], nn.newODEoptions(**kw), integrator="tsit54")
                st2 = dict(nn.ode.last_stats)
                got2 = ys[-1].local_numpy()
                ok2 = bool(np.all(np.abs(got2 - exp) <= 1e-7 * np.abs(exp))) and st2["steps"] == refl.stats.steps and st2["rejected"] == refl.stats.rejected
                if not ok2:
                    fails.append(("lorenz96 tsit54 one-kernel attempt", nl, peer_halo, st2, refl.stats.steps, float(np.max(np.abs(got2 - exp)))))
                if rank == 0:
                    print(f"[multi-gpu world={world}] lorenz96 n={nl} tsit54 one-kernel attempt (peer_halo={peer_halo}, p2p={ctx.get('p2p')}): steps={st2['steps']} "
                          f"rejected={st2['rejected']} launches={st2['launches']} (default path {st['launches']}) collectives={st2['collectives']} "
                          f"(default {st['collectives']}) ok={ok2}", flush=True)
            # RK4 (no error norm, hence no lockstep between the ranks): always the ncclSend/ncclRecv halo
            t, ys = nn.solveODE(rhs_l, gl, [0.0, 0.05], nn.newODEoptions(dt=2e-3), integrator="rk4")
            ref4 = O.solve_vector("rk4", O.rhs_lorenz96(8.0), yl, [0.0, 0.05], O.new_options(dt=2e-3)).y[-1][lo:lo + ll]
            if not np.array_equal(ys[-1].local_numpy().view(np.uint64), np.asarray(ref4).view(np.uint64)):
                fails.append(("lorenz96 rk4 one-kernel step bitwise", nl))
        finally:
            ctx.set("fuse_stencil_attempt", 1)
            ctx.set("l96_peer_halo", 1)
    # STENCIL right-hand side given as source, sharded: the plain dydt kernel exchanges radius_left / radius_right elements per evaluation,
    # the NVRTC-compiled one-kernel attempt reads its overlap from the peer-mapped ring neighbours (or one ncclSend/ncclRecv per step)
    nl = 100003
    yl = 8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(nl) / nl)
    Kd = 1.0 + 0.5 * np.sin(2 * np.pi * 3 * np.arange(nl) / nl)
    for name, expr, rl, rr, pv, cs, fnp, y_init, rtol in (
            ("lorenz96 from source", "((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, [], [8.0],
             lambda t, y: ((np.roll(y, -1) - np.roll(y, 2)) * np.roll(y, 1) - y) + 8.0, yl, 1e-7),
            ("diffusion from source", "c0*((Y(-1) - 2.0*Y(0)) + Y(1))*p0 + c1*t", 1, 1, [Kd], [0.3, 0.05],
             lambda t, y: 0.3 * ((np.roll(y, 1) - 2.0 * y) + np.roll(y, -1)) * Kd + 0.05 * t, 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(nl) / nl), 1e-9)):
        gs = nn.newVector(y_init, ctx)
        lo, ll = gs.local_offset, gs.local_len
        srhs = nn.rhsJitStencil(expr, rl, rr, [nn.newVector(p, ctx) for p in pv], cs, ctx)
        out = gs._new_like()
        assert srhs.fn(0.3, gs._h, out._h, srhs.user) == 0
        if not np.array_equal(out.local_numpy().view(np.uint64), fnp(0.3, y_init)[lo:lo + ll].view(np.uint64)):
            fails.append((name, "rhs kernel bitwise"))
        refs = O.solve_vector("tsit54", O.rhs_callback(fnp), y_init, [0.0, 0.3], O.new_options(**kw))
        for fuse, peer_halo in ((1, 1), (1, 0), (0, 1)):
            ctx.set("fuse_stencil_attempt", fuse)
            ctx.set("l96_peer_halo", peer_halo)
            t, ys = nn.solveODE(srhs, gs, [0.0, 0.3], nn.newODEoptions(**kw), integrator="tsit54")
            st = dict(nn.ode.last_stats)
            got, exp = ys[-1].local_numpy(), refs.y[-1][lo:lo + ll]
            ok = bool(np.all(np.abs(got - exp) <= rtol * np.abs(exp) + 1e-13 * np.max(np.abs(refs.y[-1])))) and st["steps"] == refs.stats.steps and st["rejected"] == refs.stats.rejected
            if not ok:
                fails.append((name, fuse, peer_halo, st, refs.stats.steps, float(np.max(np.abs(got - exp)))))
            if rank == 0:
                print(f"[multi-gpu world={world}] {name} tsit54 fuse={fuse} peer_halo={peer_halo}: steps={st['steps']} launches={st['launches']} collectives={st['collectives']} ok={ok}", flush=True)
        ctx.set("fuse_stencil_attempt", 1)
        ctx.set("l96_peer_halo", 1)
        # RK4 of the same stencil: one kernel per step, the 4 * radius halo by one ncclSend/ncclRecv per step; bit-identical to the oracle
        t, ys = nn.solveODE(srhs, gs, [0.0, 0.02], nn.newODEoptions(dt=2e-3), integrator="rk4")
        st = dict(nn.ode.last_stats)
        ref4 = O.solve_vector("rk4", O.rhs_callback(fnp), y_init, [0.0, 0.02], O.new_options(dt=2e-3)).y[-1][lo:lo + ll]
        ok4 = np.array_equal(ys[-1].local_numpy().view(np.uint64), np.asarray(ref4).view(np.uint64))
        if not ok4:
            fails.append((name, "rk4 one-kernel step bitwise"))
        if rank == 0:
            print(f"[multi-gpu world={world}] {name} rk4 one-kernel step: steps={st['steps']} launches={st['launches']} collectives={st['collectives']} ok={ok4}", flush=True)
    # right-hand side given as SOURCE, sharded: the parameter vector shards like the state; the run-time compiled
    # attempt kernel / device loop use the same in-kernel all-reduce as the built-ins
    K = 2.0 + 3.0 * np.arange(n) / (n - 1)
    gK = nn.newVector(K, ctx)
    jrhs = nn.rhsJit("c0*y*(1.0 - y/p0) + c1*t", [gK], [0.7, 0.05])
    oj = O.rhs_callback(lambda t, y: 0.7 * y * (1.0 - y / K) + 0.05 * t)
    for method in ("dopri54", "vern65"):
        refj = O.solve_vector(method, oj, y0, [0.0, 2.0], O.new_options(**kw))
        for devloop in (1, 0):
            ctx.set("device_loop", devloop)
            t, ys = nn.solveODE(jrhs, gy0, [0.0, 2.0], nn.newODEoptions(**kw), integrator=method)
            st = dict(nn.ode.last_stats)
            got, exp = ys[-1].local_numpy(), refj.y[-1][off:off + ln]
            ok = bool(np.all(np.abs(got - exp) <= 1e-9 * np.abs(exp) + 1e-13 * np.max(np.abs(refj.y[-1])))) and \
                st["steps"] == refj.stats.steps and st["rejected"] == refj.stats.rejected and st["collectives"] == st["attempts"]
            if not ok:
                fails.append(("jit rhs", method, devloop, st, refj.stats.steps, float(np.max(np.abs(got - exp)))))
            if rank == 0:
                print(f"[multi-gpu world={world}] jit rhs {method} device_loop={devloop}: steps={st['steps']} launches={st['launches']} collectives={st['collectives']} ok={ok}", flush=True)
    ctx.set("device_loop", -1)
    # trajectory consumers, sharded: element-wise, the only scalar that crosses shards is the duplicate test's count
    rngq = np.random.default_rng(11)
    Xq = np.array([0.0, 1.0, 0.5, 1.0, 2.0, 1.5, 3.0])
    base = rngq.uniform(-1.0, 1.0, (6, n))
    Yq = np.stack([base[0], base[1], base[2], base[1], base[3], base[4], base[5]])
    dv = [nn.newVector(r, ctx) for r in Yq]
    for name, ofn in (("cumtrapz", O.cumtrapz), ("cumsimpson", O.cumsimpson)):
        got = np.array([v.local_numpy() for v in getattr(nn, name)(dv, Xq)])
        exp = ofn(Yq, Xq)[:, off:off + ln]
        ok = got.shape == exp.shape and np.array_equal(got.view(np.uint64), exp.view(np.uint64))
        if not ok:
            fails.append((name + " sharded",))
        if rank == 0:
            print(f"[multi-gpu world={world}] {name} sharded bitwise ok={ok}", flush=True)
    Ybad = Yq.copy()
    Ybad[3, n - 1] += 1.0  # lives on the LAST rank only: every rank must still see the impure duplicate
    try:
        nn.cumtrapz([nn.newVector(r, ctx) for r in Ybad], Xq)
        fails.append(("impure duplicate not detected on rank", rank))
    except ValueError:
        pass
    # sharded sum(v) goes through the same allreduce
    s = gy0.sum()
    assert abs(s - O.vector_sum(y0)) <= 1e-12 * np.abs(y0).sum(), (s, O.vector_sum(y0))
    flag = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    if fails:
        print("FAILS", rank, fails, flush=True)
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
