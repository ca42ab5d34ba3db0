#!/usr/bin/env python
"""TEST INFRASTRUCTURE: the host-emulated library with world = 2 on the CPU — two rank THREADS of this process, each with
its own b200rk context, meeting in an in-process stand-in for NCCL (fake_nccl.cpp, selected with B200RK_NCCL_LIB). Covers
the host code no single-rank run reaches: sharding, the ncclAllReduce form of the error norm, the 3-element Lorenz-96 halo
per right-hand-side evaluation, and the one-kernel Lorenz-96 attempt / RK4 step with ONE halo exchange per call
(executor.cu: exchange_attempt_halo) — every shard against the unsharded CPU oracle. CUDA IPC is not emulated, so the
peer-mailbox all-reduce and the peer-mapped halo stay GPU-only. Run as a script (tests/test_multi_rank_emulation.py):
prints one `case <name> ok=<0|1>` line per check."""
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import numpy as np

import build_emul_lib

os.environ["B200RK_NCCL_LIB"] = build_emul_lib.build_fake_nccl()
os.environ["B200RK_TEST_HOST_EMULATION"] = "1"
from numericalnim_b200 import _capi

_capi.LIB_PATH = build_emul_lib.build(sanitize=os.environ.get("B200RK_TEST_EMULATION_SANITIZE", ""))
import numericalnim_b200 as nn
import oracle as O

WORLD = int(os.environ.get("B200RK_TEST_EMUL_WORLD", "2"))   # 3: the ring neighbours are two different peers
SIZES = tuple(int(x) for x in os.environ["B200RK_TEST_EMUL_SIZES"].split(",")) if os.environ.get("B200RK_TEST_EMUL_SIZES") else \
    (1000,) if os.environ.get("B200RK_TEST_EMULATION_SANITIZE") else (1000, 2600)   # the sanitizer pass: one size (several tiles per shard at 2 ranks)
KW = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
results = {}
lock = threading.Lock()


def report(rank, name, ok):
    with lock:
        results.setdefault(name, []).append(bool(ok))


def close(got, full, lo, ll, rtol=1e-7):
    """The shard [lo, lo+ll) of the unsharded reference `full`, with DESIGN.md §5's tolerance: relative to the element plus
    1e-13 of the WHOLE vector's scale (a shard of strongly damped components is orders of magnitude below it)."""
    want = np.asarray(full)[lo:lo + ll]
    return got.shape == want.shape and bool(np.all(np.abs(got - want) <= rtol * np.abs(want) + 1e-13 * np.max(np.abs(full))))


def bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and bool(np.array_equal(a.view(np.uint64), b.view(np.uint64)))


def rank_main(rank, uid, refs):
    ctx = nn.Context(0, rank, WORLD, uid)
    p2p = ctx.get("p2p")   # 1: peer mailboxes mapped (emulated CUDA IPC: a handle is the pointer) unless B200RK_P2P=0
    if rank == 0:
        print(f"info p2p={p2p}", flush=True)
    ctx.set("device_loop", 0)   # the cooperative loop's own peer exchange spins inside a kernel team: not emulated across ranks
    try:
        # ---- element-local IVP, sharded: fused attempt + ncclAllReduce of the error norm ----
        n = 5003
        lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
        y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
        gl, gy = nn.newVector(lam, ctx), nn.newVector(y0, ctx)
        lo, ll = gy.local_offset, gy.local_len
        for fuse in (1, 0):
            ctx.set("fuse_pointwise", fuse)
            t, ys = nn.solveODE(nn.rhsDiagLinear(gl), gy, [0.0, 2.0], nn.newODEoptions(**KW), integrator="dopri54")
            report(rank, f"diag dopri54 sharded fuse_pointwise={fuse}", close(ys[-1].local_numpy(), refs["diag"].y[-1], lo, ll, 1e-9))
        ctx.set("fuse_pointwise", 1)
        # config 4's shape (Vern65, diag-linear, sharded) and fixed-step RK4 (no collective at all: shards bit-identical)
        t, ys = nn.solveODE(nn.rhsDiagLinear(gl), gy, [0.0, 2.0], nn.newODEoptions(**KW), integrator="vern65")
        st = dict(nn.ode.last_stats)
        report(rank, "diag vern65 sharded", close(ys[-1].local_numpy(), refs["vern65"].y[-1], lo, ll, 1e-9) and st["steps"] == refs["vern65"].stats.steps)
        for fuse in (1, 0):
            ctx.set("fuse_pointwise", fuse)
            c0 = ctx.stats()["collectives"]
            t, ys = nn.solveODE(nn.rhsDiagLinear(gl), gy, [0.0, 0.1], nn.newODEoptions(dt=5e-3), integrator="rk4")
            report(rank, f"diag rk4 sharded bitwise, no collective, fuse_pointwise={fuse}",
                   bits(ys[-1].local_numpy(), np.asarray(refs["rk4diag"].y[-1])[lo:lo + ll]) and ctx.stats()["collectives"] == c0)
        ctx.set("fuse_pointwise", 1)
        # ---- sum(v) and the trajectory consumers, sharded (the duplicate check is the one collective of that path) ----
        s_glob = gy.sum()
        report(rank, "sum(v) sharded", abs(s_glob - O.vector_sum(y0)) <= 1e-12 * float(np.abs(y0).sum()))
        rngq = np.random.default_rng(11)
        Xq = np.array([0.0, 1.0, 0.5, 1.0, 2.0, 1.5, 3.0])
        base = rngq.uniform(-1.0, 1.0, (6, n))
        Yq = np.stack([base[0], base[1], base[2], base[1], base[3], base[4], base[5]])
        dv = [nn.newVector(r, ctx) for r in Yq]
        for name, ofn in (("cumtrapz", O.cumtrapz), ("cumsimpson", O.cumsimpson)):
            got = np.array([v.local_numpy() for v in getattr(nn, name)(dv, Xq)])
            report(rank, f"{name} sharded bitwise", bits(got, ofn(Yq, Xq)[:, lo:lo + ll]))
        Ybad = Yq.copy()
        Ybad[3, n - 1] += 1.0   # lives on the LAST rank only: every rank must still see the impure duplicate
        try:
            nn.cumtrapz([nn.newVector(r, ctx) for r in Ybad], Xq)
            report(rank, "impure duplicate on one rank raises on every rank", False)
        except ValueError:
            report(rank, "impure duplicate on one rank raises on every rank", True)
        # ---- Lorenz-96, sharded ----
        for nl in SIZES:
            yl = 8.0 + 0.5 * np.sin(2 * np.pi * 37 * np.arange(nl) / nl)
            g = nn.newVector(yl, ctx)
            lo, ll = g.local_offset, g.local_len
            rhs = nn.rhsLorenz96(8.0, ctx)
            out = g._new_like()
            rc = rhs.fn(0.0, g._h, out._h, rhs.user)
            report(rank, f"lorenz96 rhs bitwise n={nl}", rc == 0 and bits(out.local_numpy(), O.rhs_eval(O.rhs_lorenz96(8.0), 0.0, yl)[lo:lo + ll]))
            ref = refs["l96", nl]
            modes = [(0, 1, "3-element halo per evaluation"), (1, 0, "one-kernel attempt, one halo exchange per call")]
            if p2p:
                modes.append((1, 1, "one-kernel attempt, halo read in place from the peer-mapped neighbour"))
            for knob, peer_halo, what in modes:
                ctx.set("fuse_stencil_attempt", knob)
                ctx.set("l96_peer_halo", peer_halo)
                c0 = ctx.stats()["collectives"]
                t, ys = nn.solveODE(rhs, g, [0.0, 0.3], nn.newODEoptions(**KW), integrator="tsit54")
                st = dict(nn.ode.last_stats)
                ok = close(ys[-1].local_numpy(), ref.y[-1], lo, ll) and st["steps"] == ref.stats.steps and st["rejected"] == ref.stats.rejected
                report(rank, f"lorenz96 tsit54 n={nl}: {what}", ok)
                if rank == 0:
                    print(f"info lorenz96 n={nl} fuse_stencil_attempt={knob} peer_halo={peer_halo}: steps={st['steps']} attempts={st['attempts']} "
                          f"launches={st['launches']} collectives={ctx.stats()['collectives'] - c0}", flush=True)
            ctx.set("fuse_stencil_attempt", 1)
            ctx.set("l96_peer_halo", 0)
            # the resumable solver (b200rk_solver_new / advance / free): the peer mapping lives as long as the solver, across
            # several advance calls, and is closed collectively by close(); same final state as the one-call solve
            for peer_halo in ((1, 0) if p2p else (0,)):
                ctx.set("l96_peer_halo", peer_halo)
                sol = nn.Solver("tsit54", rhs, g, 0.3, nn.newODEoptions(**KW))
                steps = 0
                while True:
                    done, fin = sol.advance(2)
                    steps += done
                    if fin:
                        break
                yend = sol.state()[3].local_numpy().copy()
                sol.close()
                report(rank, f"lorenz96 tsit54 n={nl}: resumable solver, peer_halo={peer_halo}", steps == ref.stats.steps and close(yend, ref.y[-1], lo, ll))
            ctx.set("l96_peer_halo", 0)
            # one adaptive step, bit for bit against the oracle (same dt, no rejection)
            fs = O.rhs_eval(O.rhs_lorenz96(8.0), 0.0, yl)
            gf = nn.newVector(fs, ctx)
            opts = dict(absTol=1e-2, relTol=1e-2, dtMax=1.0, dtMin=1e-8, dt=0.005)
            for method in ("dopri54", "tsit54", "vern65"):
                yn_ref, fn_ref, _, err_ref, _ = O.step_vector(method, O.rhs_lorenz96(8.0), 0.0, yl, fs, 0.005, O.new_options(**opts))
                yn, fn, dt_used, err = nn.integratorStep(method, rhs, 0.0, g, gf, 0.005, nn.newODEoptions(**opts))
                report(rank, f"lorenz96 {method} one-kernel step bitwise n={nl}",
                       bits(yn.local_numpy(), yn_ref[lo:lo + ll]) and bits(fn.local_numpy(), fn_ref[lo:lo + ll]) and abs(err - err_ref) <= 1e-12 * err_ref)
            # RK4: whole trajectory bit for bit (fixed step), halo of y only
            t, ys = nn.solveODE(rhs, g, [0.0, 0.02], nn.newODEoptions(dt=2e-3), integrator="rk4")
            report(rank, f"lorenz96 rk4 one-kernel step bitwise n={nl}", bits(ys[-1].local_numpy(), np.asarray(refs["rk4", nl].y[-1])[lo:lo + ll]))
            ctx.set("fuse_stencil_attempt", 0)
    except Exception as e:  # noqa: BLE001 — the context is left alone: vectors of this frame still refer to it
        report(rank, f"rank {rank} raised {type(e).__name__}: {e}", False)
        print(f"info rank {rank} raised {type(e).__name__}: {e}", flush=True)
    else:
        ctx.close()   # vectors of this frame outlive it: the Python mirror must not reach into a closed context


def main():
    refs = {}
    n = 5003
    lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
    y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    refs["diag"] = O.solve_vector("dopri54", O.rhs_diag_linear(lam), y0, [0.0, 2.0], O.new_options(**KW))
    refs["vern65"] = O.solve_vector("vern65", O.rhs_diag_linear(lam), y0, [0.0, 2.0], O.new_options(**KW))
    refs["rk4diag"] = O.solve_vector("rk4", O.rhs_diag_linear(lam), y0, [0.0, 0.1], O.new_options(dt=5e-3))
    for nl in SIZES:
        yl = 8.0 + 0.5 * np.sin(2 * np.pi * 37 * np.arange(nl) / nl)
        refs["l96", nl] = O.solve_vector("tsit54", O.rhs_lorenz96(8.0), yl, [0.0, 0.3], O.new_options(**KW))
        refs["rk4", nl] = O.solve_vector("rk4", O.rhs_lorenz96(8.0), yl, [0.0, 0.02], O.new_options(dt=2e-3))
    uid = nn.Context.nccl_unique_id()   # binds the stand-in NCCL on the main thread
    threads = [threading.Thread(target=rank_main, args=(r, uid, refs)) for r in range(WORLD)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=600)
    hung = any(th.is_alive() for th in threads)
    bad = 0
    for name, oks in results.items():
        ok = len(oks) == WORLD and all(oks)
        bad += not ok
        print(f"case {name} ok={int(ok)}", flush=True)
    print(f"cases={len(results)} failures={bad} hung={int(hung)}", flush=True)
    os._exit(1 if (bad or hung or not results) else 0)


if __name__ == "__main__":
    main()
