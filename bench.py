#!/usr/bin/env python
"""bench.py — RK steps/sec and stage-combine HBM GB/s (fp64) of the explicit-RK hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores
    python bench.py --sweep [--out FILE]                     # BASELINE.json config 5: kernel bandwidth sweep
    python bench.py --quad [--out FILE]                      # trajectory consumers (cumtrapz / cumsimpson / hermiteInterpolate)

A "step" is one accepted Runge–Kutta step of the adaptive driver loop (ode.nim:511-541). On the general
pipeline (any user right-hand side) an attempt is S-1 right-hand-side launches, S-1 fused stage-accumulate
launches, one fused combine+error-norm launch and an 8-byte read-back; for the library's element-local built-in
right-hand sides the whole loop runs inside one persistent kernel. When sharded, the scalar
error norm is all-reduced once per attempt inside the reducing kernel (NVLink peer mailboxes).
The JSON line reports the library's default path as `value` and the general pipeline under `pipeline`.

Workload at every N: BASELINE.json configs[1] per GPU — DOPRI54, diag-linear IVP y' = -lambda .* y,
2^23 fp64 state elements PER GPU (weak scaling: N_global = n_gpus * 2^23, contiguous shards), absTol =
relTol = 1e-6, dtMax = 1, dtMin = 1e-8. `value` counts shard-steps: K accepted steps x n_gpus shards / time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200RK_JIT_CACHE", os.path.join(ROOT, ".jitcache"))  # compiled user right-hand sides (csrc/jit.cu)

SHARD_LOG2 = 23
OPTS = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
WORKLOADS = {
    # name: (integrator, rhs, log2 elements per GPU)
    "cfg2_dopri54_diag_8M": ("dopri54", "diag", 23),
    "cfg3_tsit54_lorenz96_16M": ("tsit54", "l96", 24),
    "cfg4_vern65_diag_16M_per_gpu": ("vern65", "diag", 24),
    "rk4_diag_8M": ("rk4", "diag", 23),
}
# algorithmic bytes per element per ATTEMPT of the RK kernels (RHS excluded), zero weights skipped
ALG_BYTES_PER_ELEM = {"dopri54": 8 * (32 + 7), "tsit54": 8 * (33 + 8), "vern65": 8 * (45 + 9), "rk4": 8 * 15}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def problem_arrays(n_global: int, off: int, ln: int):
    """lambda[i] = 0.1 + 9.9 i/(N-1); y0[i] = 1 + 0.5 sin(2 pi i/N), local slice."""
    i = np.arange(off, off + ln, dtype=np.float64)
    lam = 0.1 + 9.9 * i / float(n_global - 1)
    y0 = 1.0 + 0.5 * np.sin(2.0 * np.pi * i / float(n_global))
    return lam, y0


def l96_y0(n_global: int, off: int, ln: int, F: float = 8.0):
    i = np.arange(off, off + ln, dtype=np.float64)
    return F + 0.01 * np.sin(2.0 * np.pi * 37.0 * i / float(n_global))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ======================================================================================================
# our arm
# ======================================================================================================
UNIT = "RK steps/s"   # the same string in both arms and in `e2e`, so that the driver can divide them
REPEATS = 5           # the timed region (exactly K accepted steps) is measured this many times: median + min / max


def base_config(workload: str, lg: int) -> dict:
    """`config` of the JSON line: the workload only, byte-identical between `--impl ours` and `--impl reference`
    (library knobs and notes live in the top-level `knobs` / `notes` objects of our arm)."""
    integrator, rhs_kind, _ = WORKLOADS[workload]
    return {"workload": workload, "integrator": integrator, "rhs": rhs_kind, "elems_per_gpu": 1 << lg, "options": OPTS}


def pin_to_gpu_numa_node(local_rank: int):
    """Bind this rank to the host NUMA node its GPU hangs off BEFORE any pinned buffer is allocated (first touch then
    places the staging buffers next to the GPU's PCIe root): at 8 ranks the end-to-end solves otherwise push 8 x 3 x 64 MiB
    per solve through whichever socket the processes happen to start on. Returns a note for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return {"numa_node": None, "note": "single NUMA node / not reported"}
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # noqa: BLE001 — best effort: containers may hide sysfs
        return {"numa_node": None, "note": str(e)[:120]}


def parity_check(nn, ctx, dist, rank: int, world: int) -> dict:
    """Sharded (or single-GPU) results against the CPU oracle BEFORE anything is timed, so that the scaling record carries
    parity next to throughput: the three adaptive pairs + RK4 on the diag-linear IVP and Tsit54 / Vern65 / RK4 on the
    Lorenz-96 ring (one-kernel attempt; N > 1: halo read in place from the ring neighbours over NVLink), n = 100 003
    global. The oracle (oracle/, test infrastructure) is used as the CHECKER only and is outside every timed region.
    Pass = step / rejection counts equal, states within the tolerances of DESIGN.md 5 (RK4: bit-identical)."""
    import oracle as O

    n = 100003
    kw = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
    y0 = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(n) / n)
    yl = 8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(n) / n)
    glam, gy0, gl = nn.newVector(lam, ctx), nn.newVector(y0, ctx), nn.newVector(yl, ctx)
    off, ln = gy0.local_offset, gy0.local_len
    cases, max_rel, bad = [], 0.0, []

    def one(name, method, rhs, orhs, g0, h0, ts, rtol, bitwise=False, **okw):
        nonlocal max_rel
        o = dict(kw, **okw)
        ref = O.solve_vector(method, orhs, h0, ts, O.new_options(**o))
        _, ys = nn.solveODE(rhs, g0, ts, nn.newODEoptions(**o), integrator=method)
        st = dict(nn.ode.last_stats)
        got, exp = ys[-1].local_numpy(), np.asarray(ref.y[-1])[off:off + ln]
        scale = float(np.max(np.abs(ref.y[-1])))
        rel = float(np.max(np.abs(got - exp) / (np.abs(exp) + 1e-4 * scale))) if ln else 0.0
        if bitwise:
            ok = np.array_equal(got.view(np.uint64), exp.view(np.uint64))
        else:
            ok = bool(np.all(np.abs(got - exp) <= rtol * np.abs(exp) + 1e-13 * scale))
        ok = ok and st["steps"] == ref.stats.steps and st["rejected"] == ref.stats.rejected
        max_rel = max(max_rel, rel)
        cases.append(name)
        if not ok:
            bad.append({"case": name, "rank": rank, "steps": [st["steps"], ref.stats.steps], "rejected": [st["rejected"], ref.stats.rejected], "max_rel": rel})

    d_rhs, d_or = nn.rhsDiagLinear(glam), O.rhs_diag_linear(lam)
    for m in ("dopri54", "tsit54", "vern65"):
        one("diag_linear/" + m, m, d_rhs, d_or, gy0, y0, [0.0, 2.0], 1e-9)
    one("diag_linear/rk4", "rk4", d_rhs, d_or, gy0, y0, [0.0, 0.5], 0.0, bitwise=True, dt=5e-3)
    l_rhs, l_or = nn.rhsLorenz96(8.0, ctx), O.rhs_lorenz96(8.0)
    one("lorenz96/tsit54", "tsit54", l_rhs, l_or, gl, yl, [0.0, 0.5], 1e-7)
    one("lorenz96/vern65", "vern65", l_rhs, l_or, gl, yl, [0.0, 0.5], 1e-7)
    one("lorenz96/rk4", "rk4", l_rhs, l_or, gl, yl, [0.0, 0.05], 0.0, bitwise=True, dt=2e-3)
    for v in (glam, gy0, gl):
        v.free()
    n_bad, worst = len(bad), max_rel
    if dist is not None:
        import torch
        t = torch.tensor([float(n_bad), worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_bad, worst = int(t[0].item()), float(t[1].item())
    return {"ok": n_bad == 0, "cases": len(cases), "max_rel": worst, "n_global": n, "ranks": world, "names": cases,
            "failed_on_rank0": bad or None,
            "how": "every rank compares its shard with the unsharded CPU oracle (checker only, outside the timed regions): "
                   "step and rejection counts equal; states within 1e-9 (diag-linear) / 1e-7 (Lorenz-96) relative; RK4 bit-identical"}


def run_ours(args):
    import torch

    import numericalnim_b200 as nn
    from numericalnim_b200 import _capi

    rank, local_rank, world = dist_env()
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    numa = pin_to_gpu_numa_node(local_rank)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [nn.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    ctx = nn.Context(local_rank, rank, world, nccl_id)
    nn.set_default_context(ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    peak, peak_src = peaks()
    traffic_db = {}
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            traffic_db = json.load(fh)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    parity = None if args.no_parity else parity_check(nn, ctx, dist, rank, world)

    clocks = ClockSampler(local_rank)
    clocks.__enter__()  # sampled from the warm-up through the timed regions and the end-to-end solves

    def gbs(p):
        return p["bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else 0.0

    class Problem:
        """One workload's device-resident inputs (this rank's shard)."""

        def __init__(self, workload: str, log2n: int = 0):
            self.workload = workload
            self.integrator, self.rhs_kind, lg = WORKLOADS[workload]
            self.lg = log2n or lg
            self.n_shard = 1 << self.lg
            self.n_global = self.n_shard * world
            probe = nn.GpuVector.empty(self.n_global, ctx)
            self.off, self.ln = probe.local_offset, probe.local_len
            probe.free()
            self.opts = nn.newODEoptions(**OPTS)
            self.glam, self.lam_h = None, None
            if self.rhs_kind == "diag":
                self.lam_h, self.y0_h = problem_arrays(self.n_global, self.off, self.ln)
                self.glam = nn.GpuVector.from_local(self.n_global, self.lam_h, ctx)
                self.rhs = nn.rhsDiagLinear(self.glam)
            else:
                self.y0_h = l96_y0(self.n_global, self.off, self.ln)
                self.rhs = nn.rhsLorenz96(8.0, ctx)
            self.gy0 = nn.GpuVector.from_local(self.n_global, self.y0_h, ctx)

        def free(self):
            for v in (self.glam, self.gy0):
                if v is not None:
                    v.free()

    def timed_steps(pb: Problem, fuse: int, rhs=None, repeats: int = 1) -> dict:
        """`repeats` x (fresh solver, W untimed + K timed accepted steps) with the fused paths on / off; device time (CUDA
        events on the library's stream), max over ranks per repeat; then K more steps with one event pair around every
        launch (the library's profiler) for the per-kernel roofline."""
        rhs = rhs or pb.rhs
        for k in ("fuse_pointwise", "fuse_stencil", "fuse_stencil_attempt"):
            ctx.set(k, fuse)
        ctx.set("profile", 0)
        ms_all, last = [], None
        for _ in range(repeats):
            if last is not None:
                last[0].close()
            solver = nn.Solver(pb.integrator, rhs, pb.gy0, 1e12, pb.opts)
            solver.advance(args.warmup)
            st0, cs0 = solver.stats(), ctx.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            with torch.cuda.stream(stream):
                e0.record()
            done, _ = solver.advance(args.steps)
            with torch.cuda.stream(stream):
                e1.record()
            barrier()
            assert done == args.steps, (done, args.steps)
            ms_all.append(allmax(e0.elapsed_time(e1)))
            last = (solver, st0, cs0, solver.stats(), ctx.stats())
        solver, st0, cs0, st1, cs1 = last
        t_now_, dt_next_, _, _ = solver.state()
        # The same solver continues for K more steps with one CUDA-event pair around EVERY kernel launch (on the
        # launching stream): per-kernel durations for the roofline. Kept out of the timed region above because 2 event
        # records per launch cost a few percent of a ~60 us fused step.
        ctx.set("profile", 1)
        ctx.profile_reset()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            i0.record()
        solver.advance(args.steps)
        with torch.cuda.stream(stream):
            i1.record()
        barrier()
        prof_ = ctx.profile_read()
        prof_["instrumented_ms"] = i0.elapsed_time(i1)
        ctx.set("profile", 0)
        solver.close()
        med = statistics.median(ms_all)
        return dict(ms_max=med, ms_all=ms_all, prof=prof_, attempts=st1["attempts"] - st0["attempts"],
                    rejected=st1["rejected"] - st0["rejected"], launches=cs1["launches"] - cs0["launches"],
                    collectives=cs1["collectives"] - cs0["collectives"], t=t_now_, dt_next=dt_next_)

    def stage_roofline(pb: Problem, r: dict) -> dict:
        """Roofline object of the stage / RHS / finish pipeline: the dominant kernel family is stage_kernel."""
        st, fn_, rh = r["prof"]["stage"], r["prof"]["finish"], r["prof"]["rhs"]
        kms = st["ms"] + fn_["ms"] + rh["ms"] + r["prof"]["other"]["ms"]
        a = gbs(st)
        return {"bound": "hbm", "kernel": "stage_kernel<M,W,U> (fused stage accumulate y + dt*sum(a_sj k_j); all %d launches of the timed region)" % st["launches"],
                "achieved": a, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": a / peak, "frac_of_nominal_8000": a / 8000.0,
                "traffic": traffic_db.get("stage_kernel_%s_2p%d" % (pb.integrator, pb.lg)), "launches": st["launches"],
                "avg_launch_us": 1e3 * st["ms"] / max(1, st["launches"]), "algorithmic_bytes_per_launch": st["bytes"] / max(1, st["launches"]),
                "finish_kernel": {"achieved": gbs(fn_), "frac": gbs(fn_) / peak, "launches": fn_["launches"], "avg_launch_us": 1e3 * fn_["ms"] / max(1, fn_["launches"])},
                "rhs_kernel": {"achieved": gbs(rh), "launches": rh["launches"]},
                "instrumented_ms_per_step": r["prof"]["instrumented_ms"] / args.steps,
                "kernel_time_share_of_step": kms / r["prof"]["instrumented_ms"] if r["prof"]["instrumented_ms"] > 0 else None}

    def fused_roofline(pb: Problem, r: dict) -> tuple:
        """Roofline of the library's default (fused) path of a workload: (path description, roofline object)."""
        fu = r["prof"]["fused"]
        a = gbs(fu)
        common = {"bound": "hbm", "achieved": a, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": a / peak,
                  "frac_of_nominal_8000": a / 8000.0, "launches": fu["launches"], "avg_launch_us": 1e3 * fu["ms"] / max(1, fu["launches"]),
                  "algorithmic_bytes_per_launch": fu["bytes"] / max(1, fu["launches"]),
                  "instrumented_ms_per_step": r["prof"]["instrumented_ms"] / args.steps,
                  "kernel_time_share_of_step": fu["ms"] / r["prof"]["instrumented_ms"] if r["prof"]["instrumented_ms"] > 0 else None}
        if pb.rhs_kind != "diag":
            path = "l96_attempt (built-in Lorenz-96: every attempt = ONE kernel over overlapped tiles; sharded: halo read in place from the ring neighbours)"
            common.update({"kernel": "l96_attempt_kernel<PAT,J> (whole attempt: all stages, Lorenz-96 stencil, yNew, error norm; 4 vector passes of algorithmic traffic)",
                           "traffic": traffic_db.get("l96_attempt_kernel_%s_2p%d" % (pb.integrator, pb.lg))})
            return path, common
        passes = 5
        attempts_instr = fu["bytes"] / (8.0 * pb.n_shard * passes) if pb.n_shard else 0.0
        devloop = ctx.get("device_loop") != 0 and (world == 1 or bool(ctx.get("p2p")))
        if devloop:
            path = "fused_run (element-local built-in RHS: all K steps in one persistent cooperative kernel, controller on the device)"
            kname = ("fused_run_kernel<PAT,RHS> (persistent cooperative kernel: every attempt = all stages, RHS, yNew, error norm, "
                     "grid-wide exchange of the norm, device-side controller; one launch runs the K steps)")
            per_attempt = traffic_db.get("fused_run_kernel_per_attempt_%s_2p%d" % (pb.integrator, pb.lg))
            traffic = per_attempt * attempts_instr / max(1, fu["launches"]) if per_attempt else None
        else:
            path = "fused_attempt (element-local built-in RHS)"
            kname = "fused_attempt_kernel<PAT,RHS> (whole attempt of an element-local IVP in one kernel: all stages, RHS, yNew, error norm)"
            traffic = traffic_db.get("fused_attempt_kernel_%s_2p%d" % (pb.integrator, pb.lg))
        common.update({"kernel": kname, "traffic": traffic, "attempts_in_launches": attempts_instr,
                       "us_per_attempt": 1e3 * fu["ms"] / max(1.0, attempts_instr)})
        return path, common

    def summarize(pb: Problem, r: dict, roofline: dict, path: str) -> dict:
        return {"workload": pb.workload, "value": args.steps * world / (r["ms_max"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms_max"] / args.steps,
                "attempts": r["attempts"], "rejected": r["rejected"], "attempts_per_sec": r["attempts"] * world / (r["ms_max"] * 1e-3),
                "gpu_launches": r["launches"], "collectives": r["collectives"], "t_reached": r["t"], "path": path, "roofline": roofline,
                "config": base_config(pb.workload, pb.lg)}

    # ---- headline workload: the general pipeline (what any user closure gets), then the library's default path ----------
    pb = Problem(args.workload, args.log2n)
    integrator, rhs_kind, lg, n_shard, n_global = pb.integrator, pb.rhs_kind, pb.lg, pb.n_shard, pb.n_global
    fusable = not args.no_fuse
    pipe = timed_steps(pb, 0)
    head = timed_steps(pb, 1, repeats=REPEATS) if fusable else timed_steps(pb, 0, repeats=REPEATS)
    for k in ("fuse_pointwise", "fuse_stencil", "fuse_stencil_attempt"):
        ctx.set(k, 1 if fusable else 0)
    if fusable:
        path, roofline = fused_roofline(pb, head)
    else:
        path, roofline = "stage/RHS/finish pipeline", stage_roofline(pb, head)
    ms_max, attempts, launches, t_now, dt_next = head["ms_max"], head["attempts"], head["launches"], head["t"], head["dt_next"]
    pipeline_obj = None
    if fusable:
        pipeline_obj = {"note": "same K steps with the fused paths off: the stage / RHS / finish pipeline every user-supplied right-hand side runs through",
                        "value": args.steps * world / (pipe["ms_max"] * 1e-3), "ms_per_step": pipe["ms_max"] / args.steps, "attempts": pipe["attempts"],
                        "gpu_launches": pipe["launches"],
                        "hbm_gbs_step": ALG_BYTES_PER_ELEM.get(integrator, 0) * n_shard * pipe["attempts"] / (pipe["ms_max"] * 1e-3) / 1e9,
                        "roofline": stage_roofline(pb, pipe)}

    # ---- the other single-/multi-GPU configurations of BASELINE.json on the same line: config 3 (N = 1) and config 4 (every N)
    extra = {}
    if not args.no_extra_configs and fusable:
        for key, wl, cond in (("cfg3", "cfg3_tsit54_lorenz96_16M", True), ("cfg4", "cfg4_vern65_diag_16M_per_gpu", True)):
            if not cond or wl == args.workload:
                continue
            try:
                pbx = Problem(wl)
                rx = timed_steps(pbx, 1, repeats=3)
                px, rfx = fused_roofline(pbx, rx)
                extra[key] = summarize(pbx, rx, rfx, px)
                extra[key]["repeats"] = {"n": len(rx["ms_all"]), "ms": rx["ms_all"]}
                if key == "cfg3" and not args.no_jit and world == 1:   # N = 1 only: a rank-local NVRTC failure must not desynchronise a sharded run
                    # the same Lorenz-96 ring handed over as a SOURCE expression (b200rk_jit_stencil_rhs_new): NVRTC compiles it into the
                    # generic one-kernel attempt — what a user-defined stencil closure gets instead of the built-in
                    try:
                        srhs = nn.rhsJitStencil("((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, [], [8.0], ctx)
                        rs = timed_steps(pbx, 1, srhs)
                        fu = rs["prof"]["fused"]
                        extra[key]["stencil_from_source"] = {
                            "note": "right-hand side given as the source expression ((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0 with radii 2 / 1, compiled at run time "
                                    "into the overlapped-tile attempt kernel (bit-identical to the built-in: tests/test_gpu_stencil_from_source.py)",
                            "value": args.steps * world / (rs["ms_max"] * 1e-3), "ms_per_step": rs["ms_max"] / args.steps, "attempts": rs["attempts"],
                            "gpu_launches": rs["launches"], "avg_launch_us": 1e3 * fu["ms"] / max(1, fu["launches"]),
                            "hbm_gbs": gbs(fu), "frac": gbs(fu) / peak}
                    except Exception as e:  # noqa: BLE001
                        extra[key]["stencil_from_source"] = {"error": str(e)[:300]}
                        ctx.set("profile", 0)
                pbx.free()
            except Exception as e:  # noqa: BLE001 — the headline must survive
                extra[key] = {"error": str(e)[:400]}
                ctx.set("profile", 0)

    # ---- end to end: solveODE over [0, 2] from pinned HOST buffers (H2D y0 [+ lambda], solve, D2H states) ----------------
    L = _capi.lib()
    y0_h, lam_h, ln, rhs, glam, opts = pb.y0_h, pb.lam_h, pb.ln, pb.rhs, pb.glam, pb.opts
    y0_pin = torch.from_numpy(y0_h).pin_memory()
    lam_pin = torch.from_numpy(lam_h).pin_memory() if lam_h is not None else None
    out_pin = torch.empty((2, ln), dtype=torch.float64).pin_memory()
    ts = np.array([0.0, 2.0] if rhs_kind == "diag" else [0.0, 1.0])
    t_out = np.empty(2)
    n_out = C.c_size_t(0)
    method = nn.ode.method_id(integrator)
    reps = max(1, args.e2e_reps)

    def e2e_variant(upload_params: bool) -> dict:
        steps_, ms_, h2d_, d2h_ = 0, [], 0, 0
        for rep in range(reps + 1):  # first repetition is warm-up
            st = _capi.Stats()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            with torch.cuda.stream(stream):
                a0.record()
            if upload_params and lam_pin is not None:
                _capi.check(L.b200rk_vec_upload_local_async(glam._h, lam_pin.data_ptr()), ctx.handle)
            _capi.check(L.b200rk_solve_host(ctx.handle, method, rhs.fn, rhs.user, n_global, y0_pin.data_ptr(), ts.ctypes.data, 2,
                                            C.byref(opts), t_out.ctypes.data, out_pin.data_ptr(), C.byref(n_out), C.byref(st)), ctx.handle)
            with torch.cuda.stream(stream):
                a1.record()
            barrier()
            if rep == 0:
                continue
            ms_.append(allmax(a0.elapsed_time(a1)))
            steps_ += st.steps
            h2d_ += y0_pin.numel() * 8 + (lam_pin.numel() * 8 if (upload_params and lam_pin is not None) else 0)
            d2h_ += int(n_out.value) * ln * 8
        tot = sum(ms_)
        return {"value": steps_ * world / (tot * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_ / max(1, steps_), "d2h_bytes_per_step": d2h_ / max(1, steps_),
                "steps_per_solve": steps_ / reps, "ms_per_solve": tot / reps, "ms_per_solve_all": ms_,
                "h2d_bytes_per_solve": h2d_ / reps, "d2h_bytes_per_solve": d2h_ / reps}

    e2e_full = e2e_variant(True)
    e2e_res = e2e_variant(False) if lam_pin is not None else None
    # the tStart output slot (= y0): library default (D2H on a second stream while the solve runs) vs a host-side copy of y0
    auto_host = False
    ctx.set("tstart_copy", 1)
    e2e_alt = e2e_variant(True)
    ctx.set("tstart_copy", 0)
    # PCIe roofline of the end-to-end number: the same pinned buffers copied alone, both directions
    pcie = {}
    try:
        dbuf = torch.empty(ln, dtype=torch.float64, device="cuda")
        for name, dst, src in (("h2d", dbuf, y0_pin), ("d2h", out_pin[0], dbuf)):
            best = 1e30
            for _ in range(4):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                c0.record()
                dst.copy_(src, non_blocking=True)
                c1.record()
                torch.cuda.synchronize()
                best = min(best, allmax(c0.elapsed_time(c1)))
            pcie[name + "_gbs_per_gpu"] = ln * 8 / (best * 1e-3) / 1e9
        del dbuf
        for v, key in ((e2e_full, "value"), (e2e_res, "rhs_resident")):
            if v is None:
                continue
            floor_ms = 1e3 * (v["h2d_bytes_per_solve"] / (pcie["h2d_gbs_per_gpu"] * 1e9) + v["d2h_bytes_per_solve"] / (pcie["d2h_gbs_per_gpu"] * 1e9))
            v["pcie_floor_ms_per_solve"] = floor_ms
            v["pcie_share_of_solve"] = floor_ms / v["ms_per_solve"]
    except Exception as e:  # noqa: BLE001
        pcie = {"error": str(e)[:200]}
    clocks.__exit__(None, None, None)

    # ---- the same IVP with the right-hand side handed over as SOURCE ("-(p0*y)"): NVRTC compiles it into the fused
    # kernels at run time (csrc/jit.cu) — what a user-defined element-local closure gets instead of the built-in.
    # Runs after every other GPU measurement so that nothing it does can disturb them.
    jit_obj = None
    if rhs_kind == "diag" and fusable and world == 1 and not args.no_jit:  # N = 1 only: a rank-local failure must not desynchronise a sharded run
        try:
            t0 = time.time()
            jrhs = nn.rhsJit("-(p0*y)", [glam])
            jr = timed_steps(pb, 1, jrhs)
            fu = jr["prof"]["fused"]
            jit_obj = {"note": "right-hand side given as the source expression -(p0*y), compiled at run time (NVRTC) into the same fused kernels",
                       "value": args.steps * world / (jr["ms_max"] * 1e-3), "ms_per_step": jr["ms_max"] / args.steps, "attempts": jr["attempts"],
                       "gpu_launches": jr["launches"], "t_reached": jr["t"],
                       "hbm_gbs": fu["bytes"] / (fu["ms"] * 1e-3) / 1e9 if fu["ms"] > 0 else None,
                       "wall_s_including_compile": time.time() - t0}
        except Exception as e:  # noqa: BLE001 — the headline must survive a missing NVRTC
            jit_obj = {"error": str(e)[:400]}
            try:
                ctx.set("profile", 0)
            except Exception:  # noqa: BLE001
                pass
    # ---- consumers of the trajectory (hermiteInterpolate, cumtrapz, cumsimpson; csrc/quadrature.cu): bandwidth at the
    # workload's vector length, 17 time points — `bench.py --quad` in a child process (own CUDA context on the same
    # GPU), after every solver measurement, so that nothing it does can disturb them.
    quad_obj = None
    if world == 1 and not args.no_quad:
        try:
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--quad", "--log2n", str(lg), "--quad-points", "17", "--quad-iters", "3"],
                                capture_output=True, text=True, timeout=600, env=dict(os.environ, LOCAL_RANK=str(local_rank)))
            rows = [json.loads(l) for l in cp.stdout.splitlines() if l.startswith('{"op"')]
            if cp.returncode != 0 or not rows:
                raise RuntimeError("bench.py --quad exited %d: %s" % (cp.returncode, cp.stderr.strip()[-300:]))
            quad_obj = {"note": "trajectory consumers at 2^%d elements x 17 points: GB/s of algorithmic bytes over kernel time (bench.py --quad)" % lg,
                        "rows": [{k: r[k] for k in ("op", "points", "kernel_ms_per_call", "GBps", "frac_of_peak", "cpu_GBps_same_bytes")} for r in rows]}
        except Exception as e:  # noqa: BLE001
            quad_obj = {"error": str(e)[:400]}
    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample -----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(integrator, rhs_kind, n_shard, steps=2)

    if rank == 0:
        e2e = dict(e2e_full)
        e2e["note"] = ("complete solveODE calls through b200rk_solve_host on pinned HOST buffers, inside the timed region every solve: H2D y0 and lambda "
                       "(the right-hand side's parameter vector), the solve, D2H of the returned states; accepted steps / device time (max over ranks)")
        if e2e_res is not None:
            e2e["rhs_resident"] = dict(e2e_res, note="same, with lambda left on the device between solves (it belongs to the right-hand-side object, "
                                                     "like the data a reference ODEProc closure captures): only y0 goes up, the states come down")
        e2e["tstart_slot"] = {"default": "host-side copy of y0 by helper threads" if auto_host else "device-to-host copy on a second stream while the solve runs",
                              "alternative": "device-to-host copy on a second stream" if auto_host else "host-side copy of y0 by helper threads",
                              "alternative_value": e2e_alt["value"], "alternative_ms_per_solve": e2e_alt["ms_per_solve"]}
        e2e["pcie"] = pcie
        e2e["host_numa"] = numa
        ms_all = head["ms_all"]
        line = {
            "metric": "rk_steps_per_sec", "value": args.steps * world / (ms_max * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(args.workload, lg),
            "repeats": {"n": len(ms_all), "ms_per_K_steps": ms_all, "median_ms": ms_max, "min_ms": min(ms_all), "max_ms": max(ms_all),
                        "value_from": "median", "value_at_min": args.steps * world / (min(ms_all) * 1e-3), "value_at_max": args.steps * world / (max(ms_all) * 1e-3)},
            "notes": {"value": "accepted RK steps of one 2^%d-element shard per second, summed over the %d GPU(s) (weak scaling: every rank advances its own shard "
                               "of one N_global = %d system in lockstep)" % (lg, world, n_global),
                      "elems_global": n_global,
                      "l2": "working set (>= 5 vectors x %d MiB per GPU) exceeds the 126 MB L2; no flush" % (n_shard * 8 >> 20),
                      "sharding": "contiguous shards, 1 all-reduce(sum, 1 x f64) of the error norm per attempt" if world > 1 else "single GPU, no collective",
                      "error_norm_allreduce": ("in-kernel peer mailboxes over NVLink (CUDA IPC)" if ctx.get("p2p") else "ncclAllReduce") if world > 1 else None,
                      "butcher_row": "kernel parameters / constant bank (uniform broadcast) instead of shared-memory staging (DESIGN.md 4)"},
            "knobs": {k: ctx.get(k) for k in ("vec_width", "ctas_per_sm", "finish_ctas_per_sm", "fuse_pointwise", "fuse_stencil", "fuse_stencil_attempt",
                                              "fused_ctas_per_sm", "spin_readback", "l2_hints", "device_loop", "tstart_copy", "peer_timeout_s")},
            "attempts": attempts, "attempts_per_sec": attempts * world / (ms_max * 1e-3), "rejected": head["rejected"],
            "t_reached": t_now, "dt_next": dt_next,
            "gpu_launches": launches, "collectives": head["collectives"],
            "path": path,
            "roofline": roofline,
            "parity_check": parity,
            "pipeline": pipeline_obj,
            "cfg3": extra.get("cfg3"),
            "cfg4": extra.get("cfg4"),
            "jit_rhs": jit_obj,
            "trajectory_consumers": quad_obj,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ======================================================================================================
# CPU: oracle port (the reference cannot be built: Nim, no toolchain) — bench.py's only use of oracle/
# ======================================================================================================
def oracle_problem(O, rhs_kind: str, n: int):
    if rhs_kind == "diag":
        lam, y0 = problem_arrays(n, 0, n)
        return O.rhs_diag_linear(lam), y0, lam
    return O.rhs_lorenz96(8.0), l96_y0(n, 0, n), None


def cpu_baseline_sample(integrator: str, rhs_kind: str, n: int, steps: int):
    import oracle as O

    orhs, y0, lam = oracle_problem(O, rhs_kind, n)
    t0 = time.time()
    s = O.solve_vector(integrator, orhs, y0, [0.0, 1e12], O.new_options(**OPTS), max_steps=steps)
    wall = time.time() - t0
    out = {"value": s.stats.steps / s.stats.seconds, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": "%d accepted %s steps (+2 start-up RHS evaluations) at the full N=2^%d on 1 of %d host cores (numericalnim is single-threaded); %.1f s" % (
               s.stats.steps, integrator, int(np.log2(n)), os.cpu_count() or 0, wall)}
    if lam is not None:
        try:  # courtesy row: NOT the reference's behaviour — the attempt fused into one pass, all host cores
            _, fs = O.fused_mt_solve_diag(integrator, lam, y0, 1e12, O.new_options(**OPTS), max_steps=10)
            out["courtesy_fused_multithreaded"] = {"value": fs.steps / fs.seconds, "unit": UNIT, "cores": int(fs.threads),
                                                   "note": "same IVP, whole attempt fused into one pass per element (5 vector passes), std::thread over all host cores, scalar -O2 code; "
                                                           "element-wise results bit-identical to the port. numericalnim itself is single-threaded and allocates a Vector per operator."}
        except Exception as e:  # noqa: BLE001
            out["courtesy_fused_multithreaded"] = {"error": str(e)[:200]}
    return out


def run_reference(args):
    """The reference's own CPU implementation of the path, as faithfully as it can be had here: the C++
    oracle port (allocating one-pass-per-operator Vectors, sequential sum, -O2, no FMA). Single thread because
    numericalnim's ode.nim / utils.nim have no threading construct — that is every thread the reference uses.
    Runs the FULL workload size (2^23 elements: ~6.5 s per step); only when W + K steps would not fit --cpu-budget-s does
    it fall back to a power-of-two sample, scaling the time linearly in N (which favours the CPU) and saying so."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import oracle as O

    integrator, rhs_kind, lg = WORKLOADS[args.workload]
    lg = args.log2n or lg
    n_full = 1 << lg
    warm = min(args.warmup, 1)   # a CPU has no clocks to ramp or JIT to warm: one untimed step pages the buffers in
    # calibrate on ONE step at the full size (page faults and all), then pick the largest power-of-two N_s that fits
    orhs, y0, _ = oracle_problem(O, rhs_kind, n_full)
    c = O.solve_vector(integrator, orhs, y0, [0.0, 1e12], O.new_options(**OPTS), max_steps=1)
    per_step_full = c.stats.seconds
    n_s = n_full
    while n_s > (1 << 16) and per_step_full * (n_s / n_full) * (args.steps + warm) > args.cpu_budget_s:
        n_s >>= 1
    if n_s != n_full:
        orhs, y0, _ = oracle_problem(O, rhs_kind, n_s)
    if warm and n_s != n_full:
        O.solve_vector(integrator, orhs, y0, [0.0, 1e12], O.new_options(**OPTS), max_steps=warm)
    t0 = time.time()
    s = O.solve_vector(integrator, orhs, y0, [0.0, 1e12], O.new_options(**OPTS), max_steps=args.steps)
    wall = time.time() - t0
    scale = n_full / n_s  # time assumed linear in N (bandwidth-bound) — favours the CPU, see DESIGN.md 6
    secs = s.stats.seconds * scale
    value = s.stats.steps / secs
    sample = "%d accepted %s steps at N_s=2^%d (%s of the 2^%d workload%s), 1 of %d host cores (the reference is single-threaded), %.1f s wall" % (
        s.stats.steps, integrator, int(np.log2(n_s)), "all" if n_s == n_full else "1/%d" % int(scale), lg,
        "" if n_s == n_full else ", time scaled x%g" % scale, os.cpu_count() or 0, wall)
    line = {"impl": "reference", "metric": "rk_steps_per_sec", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, s.stats.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(args.workload, lg),
            "notes": {"value": "CPU arm: one 2^%d-element shard on one host thread (the reference is single-threaded); not multiplied by n_gpus" % lg,
                      "warmup": "%d untimed step(s) at the full size (the calibration step), then exactly K timed steps" % 1},
            "knobs": {},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ======================================================================================================
# config 5: kernel bandwidth sweep
# ======================================================================================================
def run_sweep(args):
    """BASELINE.json config 5: fused stage-combine / combine+norm bandwidth and solver steps/s for N = 2^min..2^max
    (N is the GLOBAL length; under torchrun it is sharded over the ranks), next to the oracle port's CPU time for the
    same expressions on one host core (rank 0, N <= 2^cpu_max)."""
    import torch

    import numericalnim_b200 as nn
    import oracle as O
    from numericalnim_b200 import distributed as D

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = D.init_context(local_rank)
    nn.set_default_context(ctx)
    peak, peak_src = peaks()
    rows = []
    l2_bytes = 126e6
    rng = np.random.default_rng(1234)
    w5 = O.pair_tableau("dopri54")["a"][6][:5]
    w8 = O.pair_tableau("vern65")["a"][9][:8]

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for lg in range(args.sweep_min, args.sweep_max + 1, args.sweep_step):
        n = 1 << lg
        off, ln = D.shard_range(n, rank, world)
        base = rng.uniform(-1.0, 1.0, ln)
        vecs = [nn.GpuVector.from_local(n, np.roll(base, 17 * j), ctx) for j in range(10)]
        out = nn.GpuVector.empty(n, ctx)
        cases = []
        for m, w in ((1, w5), (5, w5), (8, w8)):
            cases.append((f"stage_m{m}", "stage", lambda m=m, w=w: nn.stageAccum(w[:m], 0.01, vecs[0], vecs[1:1 + m], out=out)))
        for meth, S in (("dopri54", 7), ("tsit54", 7), ("vern65", 9)):
            cases.append((f"finish_{meth}", "finish", lambda meth=meth, S=S: nn.combineErr(meth, 0.01, 1e-6, 1e-6, vecs[0], vecs[1:1 + S])))
        for name, cls, fn in cases:
            for _ in range(args.sweep_warm):
                fn()
            ctx.set("profile", 1)
            ctx.profile_reset()
            for _ in range(args.sweep_iters):
                fn()
            p = ctx.profile_read()
            ctx.set("profile", 0)
            ms, b, nl = allmax(p[cls]["ms"]), p[cls]["bytes"], p[cls]["launches"]
            gbs = world * b / (ms * 1e-3) / 1e9  # every rank moves the same bytes; slowest rank's time
            row = {"log2n": lg, "n_gpus": world, "kernel": name, "us_per_launch": 1e3 * ms / nl, "GBps": gbs, "GBps_per_gpu": gbs / world,
                   "frac_of_peak_per_gpu": gbs / world / peak, "l2_resident": bool(b / nl < l2_bytes), "alg_bytes_per_launch_per_gpu": b / nl}
            if rank == 0 and world == 1 and lg <= args.sweep_cpu_max and name in ("stage_m5", "finish_dopri54"):
                hv = [v.local_numpy() for v in vecs[:8]]
                best = 1e30
                for _ in range(3):
                    t0 = time.perf_counter()
                    if name == "stage_m5":
                        O.weighted_stage(w5, 0.01, hv[0], hv[1:6])
                    else:
                        O.pair_finish("dopri54", 0.01, 1e-6, 1e-6, hv[0], hv[1:8])
                    best = min(best, time.perf_counter() - t0)
                # same algorithmic-byte formula as the GPU column (the reference really moves ~5-6x more)
                row["cpu_ms"] = 1e3 * best
                row["cpu_GBps"] = (b / nl) / best / 1e9
                row["gpu_over_cpu"] = best / (1e-3 * ms / nl)
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
        # solver steps/s at this N (DOPRI54, diag-linear), pipeline and fused element-local path
        i = np.arange(off, off + ln, dtype=np.float64)
        glam = nn.GpuVector.from_local(n, 0.1 + 9.9 * i / float(max(n - 1, 1)), ctx)
        gy0 = nn.GpuVector.from_local(n, 1.0 + 0.5 * np.sin(2.0 * np.pi * i / float(n)), ctx)
        rhs = nn.rhsDiagLinear(glam)
        # general pipeline / fused attempt driven by the host / fused attempt inside the persistent device loop
        for name, fuse, devloop in (("pipeline", 0, 0), ("fused", 1, 0), ("fused_device_loop", 1, 1)):
            if devloop and world > 1 and not ctx.get("p2p"):
                continue
            ctx.set("fuse_pointwise", fuse)
            ctx.set("device_loop", devloop)
            sv = nn.Solver("dopri54", rhs, gy0, 1e12, nn.newODEoptions(**OPTS))
            sv.advance(8)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sv.advance(args.sweep_solver_steps)
            torch.cuda.synchronize()
            dt_s = allmax(time.perf_counter() - t0)
            sv.close()
            row = {"log2n": lg, "n_gpus": world, "kernel": "solver_dopri54_" + name,
                   "steps_per_sec": args.sweep_solver_steps / dt_s, "us_per_step": 1e6 * dt_s / args.sweep_solver_steps}
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
        ctx.set("device_loop", -1)
        ctx.set("fuse_pointwise", 1)
        for v in vecs + [out, glam, gy0]:
            v.free()
        ctx.set("pool_budget_mb", 0)
        ctx.set("pool_budget_mb", 48 << 10)
    if args.out and rank == 0:
        with open(args.out, "w") as fh:
            json.dump({"peak_gbs": peak, "peak_source": peak_src, "n_gpus": world, "host_cores": os.cpu_count(), "rows": rows}, fh, indent=1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def quad_rows(nn, O, ctx, lg: int, m: int, iters: int, cpu_log2n: int, peak: float, emit=None) -> list:
    """Time cumtrapz / cumsimpson over a sampled trajectory of m device vectors and hermiteInterpolate at 2m sample
    points, 2^lg elements per vector. GB/s = algorithmic bytes (every input vector read once, every output written
    once) / CUDA-event time of the kernels (library profiler, class "quad"), next to the oracle port's time for the
    same call on one host core at 2^cpu_log2n elements."""
    n = 1 << lg
    rng = np.random.default_rng(1234)
    base = rng.uniform(-1.0, 1.0, n)
    X = np.linspace(0.0, 2.0, m) + rng.uniform(-0.02, 0.02, m) * (np.arange(m) % 2)   # uneven spacing, sorted
    Y = [nn.newVector(np.roll(base, 17 * k), ctx) for k in range(m)]
    dY = [nn.newVector(np.roll(base, 17 * k + 5), ctx) for k in range(m)]
    xs = np.sort(rng.uniform(X[0], X[-1], 2 * m))
    n_cpu = 1 << min(lg, cpu_log2n)
    Yc = np.stack([np.roll(base[:n_cpu], 17 * k) for k in range(m)])
    dYc = np.stack([np.roll(base[:n_cpu], 17 * k + 5) for k in range(m)])
    cases = (("cumtrapz", lambda: nn.cumtrapz(Y, X), lambda: O.cumtrapz(Yc, X)),
             ("cumsimpson", lambda: nn.cumsimpson(Y, X), lambda: O.cumsimpson(Yc, X)),
             ("hermiteInterpolate", lambda: nn.hermiteInterpolate(xs, X, Y, dY), lambda: O.hermite_interpolate(xs, X, Yc, dYc)))
    rows = []
    try:
        for name, gpu, cpu in cases:
            for _ in range(3):
                for v in gpu():
                    v.free()
            ctx.set("profile", 1)
            ctx.profile_reset()
            t0 = time.perf_counter()
            for _ in range(iters):
                for v in gpu():
                    v.free()
            ctx.synchronize()
            wall = (time.perf_counter() - t0) / iters
            p = ctx.profile_read()["quad"]
            ctx.set("profile", 0)
            t0 = time.perf_counter()
            cpu()
            cpu_s = time.perf_counter() - t0
            gbs = p["bytes"] / (p["ms"] * 1e-3) / 1e9
            alg_per_call = p["bytes"] / iters
            row = {"op": name, "log2n": lg, "points": m, "kernel_launches_per_call": p["launches"] / iters, "kernel_ms_per_call": p["ms"] / iters,
                   "wall_ms_per_call": 1e3 * wall, "algorithmic_GB_per_call": alg_per_call / 1e9, "GBps": gbs, "frac_of_peak": gbs / peak,
                   "cpu_oracle_s_at_2p%d" % int(np.log2(n_cpu)): cpu_s, "cpu_GBps_same_bytes": alg_per_call * (n_cpu / n) / cpu_s / 1e9,
                   "gpu_over_cpu_per_element": (cpu_s / n_cpu) / (1e-3 * wall / n)}
            rows.append(row)
            if emit:
                emit(row)
    finally:
        ctx.set("profile", 0)
        for v in Y + dY:
            v.free()
    return rows


def run_quad(args):
    """SURVEY.md §8f rank 4 — bandwidth of the trajectory consumers (csrc/quadrature.cu), one JSON row per routine."""
    import numericalnim_b200 as nn
    import oracle as O

    ctx = nn.default_context()
    peak, peak_src = peaks()
    rows = quad_rows(nn, O, ctx, args.log2n or 23, args.quad_points, args.quad_iters, args.quad_cpu_log2n, peak,
                     emit=lambda row: print(json.dumps(row), flush=True))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump({"peak_gbs": peak, "peak_source": peak_src, "host_cores": os.cpu_count(), "rows": rows}, fh, indent=1)


def run_tune(args):
    """Launch-geometry matrix at one size: vec_width x ctas_per_sm for stage m=1,5,8 and the DOPRI54 finish."""
    import numericalnim_b200 as nn
    import oracle as O

    ctx = nn.default_context()
    peak, _ = peaks()
    n = 1 << (args.log2n or 23)
    rng = np.random.default_rng(1234)
    base = rng.uniform(-1.0, 1.0, n)
    vecs = [nn.newVector(np.roll(base, 17 * j), ctx) for j in range(10)]
    out = nn.GpuVector.empty(n, ctx)
    w8 = O.pair_tableau("vern65")["a"][9][:8]
    for vw in (2, 4):
        for cps in (0, 1, 2, 3, 4, 8):
            ctx.set("vec_width", vw)
            ctx.set("ctas_per_sm", cps)
            ctx.set("finish_ctas_per_sm", cps)
            res = {}
            for name, cls, fn in (("stage_m1", "stage", lambda: nn.stageAccum(w8[:1], 0.01, vecs[0], vecs[1:2], out=out)),
                                  ("stage_m5", "stage", lambda: nn.stageAccum(w8[:5], 0.01, vecs[0], vecs[1:6], out=out)),
                                  ("stage_m8", "stage", lambda: nn.stageAccum(w8, 0.01, vecs[0], vecs[1:9], out=out)),
                                  ("finish_dopri54", "finish", lambda: nn.combineErr("dopri54", 0.01, 1e-6, 1e-6, vecs[0], vecs[1:8])),
                                  ("finish_vern65", "finish", lambda: nn.combineErr("vern65", 0.01, 1e-6, 1e-6, vecs[0], vecs[1:10]))):
                for _ in range(10):
                    fn()
                ctx.set("profile", 1)
                ctx.profile_reset()
                for _ in range(50):
                    fn()
                p = ctx.profile_read()
                ctx.set("profile", 0)
                res[name] = round(p[cls]["bytes"] / (p[cls]["ms"] * 1e-3) / 1e9, 1)
            print(json.dumps({"log2n": int(np.log2(n)), "vec_width": vw, "ctas_per_sm": cps, "GBps": res,
                              "frac": {k: round(v / peak, 3) for k, v in res.items()}}), flush=True)
    # fused attempt kernel (element-local RHS): persistent-grid width x vector width
    ctx.set("ctas_per_sm", 0)
    ctx.set("finish_ctas_per_sm", 2)
    ctx.set("fuse_pointwise", 1)
    lam = nn.newVector(0.1 + 9.9 * np.arange(n) / (n - 1), ctx)
    rhs = nn.rhsDiagLinear(lam)
    o = nn.newODEoptions(absTol=1e-3, relTol=1e-3, dtMax=1.0, dtMin=1e-8)
    for vw in (2, 4):
        for cps in (1, 2, 3, 4, 6, 8, 0):
            ctx.set("vec_width", vw)
            ctx.set("fused_ctas_per_sm", cps)
            res = {}
            for meth in ("dopri54", "tsit54", "vern65", "rk4"):
                for _ in range(5):
                    nn.integratorStep(meth, rhs, 0.0, vecs[0], vecs[1], 1e-3, o)
                ctx.set("profile", 1)
                ctx.profile_reset()
                for _ in range(30):
                    nn.integratorStep(meth, rhs, 0.0, vecs[0], vecs[1], 1e-3, o)
                p = ctx.profile_read()
                ctx.set("profile", 0)
                res[meth] = {"GBps": round(p["fused"]["bytes"] / (p["fused"]["ms"] * 1e-3) / 1e9, 1), "us": round(1e3 * p["fused"]["ms"] / p["fused"]["launches"], 1)}
            print(json.dumps({"log2n": int(np.log2(n)), "kernel": "fused_attempt", "vec_width": vw, "fused_ctas_per_sm": cps, "res": res}), flush=True)
    ctx.set("vec_width", 4)
    ctx.set("fused_ctas_per_sm", 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_dopri54_diag_8M", choices=sorted(WORKLOADS))
    ap.add_argument("--log2n", type=int, default=0, help="override log2 of elements per GPU")
    ap.add_argument("--e2e-reps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-quad", action="store_true", help="skip the trajectory-consumer bandwidth leg")
    ap.add_argument("--no-jit", action="store_true", help="skip the run-time compiled right-hand-side leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded parity check against the oracle that precedes the timing")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the cfg3 / cfg4 objects of the default line")
    ap.add_argument("--no-fuse", action="store_true", help="headline = the stage/RHS/finish pipeline even for element-local built-in RHS")
    ap.add_argument("--cpu-budget-s", type=float, default=280.0)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--quad", action="store_true", help="trajectory consumers: cumtrapz / cumsimpson / hermiteInterpolate bandwidth")
    ap.add_argument("--quad-points", type=int, default=33)
    ap.add_argument("--quad-iters", type=int, default=5)
    ap.add_argument("--quad-cpu-log2n", type=int, default=18)
    ap.add_argument("--tune", action="store_true")
    ap.add_argument("--sweep-min", type=int, default=16)
    ap.add_argument("--sweep-max", type=int, default=26)
    ap.add_argument("--sweep-step", type=int, default=1)
    ap.add_argument("--sweep-warm", type=int, default=20)
    ap.add_argument("--sweep-iters", type=int, default=100)
    ap.add_argument("--sweep-cpu-max", type=int, default=23)
    ap.add_argument("--sweep-solver-steps", type=int, default=40)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.tune:
        return run_tune(args)
    if args.quad:
        return run_quad(args)
    if args.sweep:
        return run_sweep(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
