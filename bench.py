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
def run_ours(args):
    import torch

    import numericalnim_b200 as nn
    from numericalnim_b200 import _capi

    rank, local_rank, world = dist_env()
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
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
    integrator, rhs_kind, lg = WORKLOADS[args.workload]
    if args.log2n:
        lg = args.log2n
    n_shard = 1 << lg
    n_global = n_shard * world
    probe = nn.GpuVector.empty(n_global, ctx)
    off, ln = probe.local_offset, probe.local_len
    probe.free()
    opts = nn.newODEoptions(**OPTS)
    if rhs_kind == "diag":
        lam_h, y0_h = problem_arrays(n_global, off, ln)
        glam = nn.GpuVector.from_local(n_global, lam_h, ctx)
        rhs = nn.rhsDiagLinear(glam)
    else:
        y0_h = l96_y0(n_global, off, ln)  # sharded: 3-element ring halo per RHS evaluation (ncclSend/Recv)
        lam_h = None
        rhs = nn.rhsLorenz96(8.0, ctx)
    gy0 = nn.GpuVector.from_local(n_global, y0_h, ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region: exactly K accepted steps of the driver loop --------------------
    clocks = ClockSampler(local_rank)
    clocks.__enter__()  # sampled from the warm-up through the timed regions and the end-to-end solves

    def timed_steps(fuse: int, rhs=rhs) -> dict:
        """W untimed + K timed accepted steps with the given fuse_pointwise setting; device time (CUDA events on
        the library stream), max over ranks; per-kernel-class event times from the library's profiler."""
        ctx.set("fuse_pointwise", fuse)
        ctx.set("fuse_stencil", fuse)
        ctx.set("profile", 0)
        solver = nn.Solver(integrator, rhs, gy0, 1e12, opts)
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
        ms_loc = e0.elapsed_time(e1)
        st1, cs1 = solver.stats(), ctx.stats()
        t_now_, dt_next_, _, _ = solver.state()
        assert done == args.steps, (done, args.steps)
        # The same solver continues for K more steps with one CUDA-event pair around EVERY kernel launch (the
        # library's profiler, on the launching stream): per-kernel durations for the roofline. Kept out of the
        # region above because 2 event records per launch cost a few percent of a ~80 us fused step.
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
        ms_t = torch.tensor([ms_loc], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        return dict(ms=ms_loc, ms_max=float(ms_t.item()), prof=prof_, attempts=st1["attempts"] - st0["attempts"],
                    rejected=st1["rejected"] - st0["rejected"], launches=cs1["launches"] - cs0["launches"],
                    collectives=cs1["collectives"] - cs0["collectives"], t=t_now_, dt_next=dt_next_)

    # fused paths for built-in right-hand sides: diag-linear (element-local) -> whole attempt in one kernel;
    # Lorenz-96 on one GPU -> stage accumulate + stencil in one kernel (shared-memory tile)
    fusable = not args.no_fuse and (rhs_kind == "diag" or world == 1)
    pipe = timed_steps(0)                       # stage / RHS / finish pipeline (what any user closure gets)
    head = timed_steps(1) if fusable else pipe  # headline: the library's default path for this workload
    # experimental (knob fuse_stencil_attempt, off by default): the whole Lorenz-96 attempt in ONE kernel — reported as
    # an extra object next to the default path, on request
    l96_attempt = None
    if args.l96_attempt and rhs_kind != "diag" and not args.no_fuse:   # sharded too: one 12 + 8 element halo exchange per step
        try:
            ctx.set("fuse_stencil_attempt", 1)
            l96_attempt = timed_steps(1)
        except Exception as e:  # noqa: BLE001 — the headline must survive an experimental path
            l96_attempt = {"error": str(e)[:400]}
        finally:
            ctx.set("fuse_stencil_attempt", 0)
            ctx.set("profile", 0)
    ctx.set("fuse_pointwise", 1 if fusable else 0)
    ctx.set("fuse_stencil", 1 if fusable else 0)
    ms, ms_max, prof = head["ms"], head["ms_max"], head["prof"]
    attempts, launches, t_now, dt_next = head["attempts"], head["launches"], head["t"], head["dt_next"]

    # ---- end to end: solveODE over [0, 2] from pinned HOST buffers (H2D y0 + lambda, solve, D2H states) ----
    L = _capi.lib()
    y0_pin = torch.from_numpy(y0_h).pin_memory()
    lam_pin = torch.from_numpy(lam_h).pin_memory() if lam_h is not None else None
    out_pin = torch.empty((2, ln), dtype=torch.float64).pin_memory()
    ts = np.array([0.0, 2.0] if rhs_kind == "diag" else [0.0, 1.0])
    t_out = np.empty(2)
    n_out = C.c_size_t(0)
    method = nn.ode.method_id(integrator)
    e2e_steps, e2e_ms, h2d, d2h = 0, 0.0, 0, 0
    reps = max(1, args.e2e_reps)
    for rep in range(reps + 1):  # first repetition is warm-up
        st = _capi.Stats()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            a0.record()
        if lam_pin is not None:
            _capi.check(L.b200rk_vec_upload_local(glam._h, lam_pin.data_ptr()), ctx.handle)
        _capi.check(L.b200rk_solve_host(ctx.handle, method, rhs.fn, rhs.user, n_global, y0_pin.data_ptr(), ts.ctypes.data, 2,
                                        C.byref(opts), t_out.ctypes.data, out_pin.data_ptr(), C.byref(n_out), C.byref(st)), ctx.handle)
        with torch.cuda.stream(stream):
            a1.record()
        barrier()
        if rep == 0:
            continue
        e2e_ms += a0.elapsed_time(a1)
        e2e_steps += st.steps
        h2d += y0_pin.numel() * 8 + (lam_pin.numel() * 8 if lam_pin is not None else 0)
        d2h += int(n_out.value) * ln * 8
    clocks.__exit__(None, None, None)
    e2e_t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms_max = float(e2e_t.item())

    # ---- the same IVP with the right-hand side handed over as SOURCE ("-(p0*y)"): NVRTC compiles it into the fused
    # kernels at run time (csrc/jit.cu) — what a user-defined element-local closure gets instead of the built-in.
    # Runs after every other GPU measurement so that nothing it does can disturb them.
    jit_obj = None
    if rhs_kind == "diag" and fusable and world == 1 and not args.no_jit:  # N = 1 only: a rank-local failure must not desynchronise a sharded run
        try:
            t0 = time.time()
            jrhs = nn.rhsJit("-(p0*y)", [glam])
            jr = timed_steps(1, jrhs)
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
    if rank == 0 and world == 1 and not args.no_cpu_baseline and rhs_kind == "diag":
        cpu = cpu_baseline_sample(integrator, n_shard, steps=2)

    if rank == 0:
        peak, peak_src = peaks()
        traffic_db = {}
        tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(tp):
            with open(tp) as fh:
                traffic_db = json.load(fh)

        def gbs(p):
            return p["bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else 0.0

        def stage_roofline(r):
            """Roofline object of the stage / RHS / finish pipeline: the dominant kernel family is stage_kernel."""
            st, fn_, rh = r["prof"]["stage"], r["prof"]["finish"], r["prof"]["rhs"]
            kms = st["ms"] + fn_["ms"] + rh["ms"] + r["prof"]["other"]["ms"]
            a = gbs(st)
            return {"bound": "hbm", "kernel": "stage_kernel<M,W,U> (fused stage accumulate y + dt*sum(a_sj k_j); all %d launches of the timed region)" % st["launches"],
                    "achieved": a, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": a / peak, "frac_of_nominal_8000": a / 8000.0,
                    "traffic": traffic_db.get("stage_kernel_%s_2p%d" % (integrator, lg)), "launches": st["launches"],
                    "avg_launch_us": 1e3 * st["ms"] / max(1, st["launches"]), "algorithmic_bytes_per_launch": st["bytes"] / max(1, st["launches"]),
                    "finish_kernel": {"achieved": gbs(fn_), "frac": gbs(fn_) / peak, "launches": fn_["launches"], "avg_launch_us": 1e3 * fn_["ms"] / max(1, fn_["launches"])},
                    "rhs_kernel": {"achieved": gbs(rh), "launches": rh["launches"]},
                    "instrumented_ms_per_step": r["prof"]["instrumented_ms"] / args.steps,
                    "kernel_time_share_of_step": kms / r["prof"]["instrumented_ms"] if r["prof"]["instrumented_ms"] > 0 else None}

        pipeline_obj = None
        if head is not pipe:
            pipeline_obj = {"note": "same K steps with the fused paths off: the stage / RHS / finish pipeline every user-supplied right-hand side runs through",
                            "value": args.steps * world / (pipe["ms_max"] * 1e-3), "ms_per_step": pipe["ms_max"] / args.steps, "attempts": pipe["attempts"],
                            "gpu_launches": pipe["launches"],
                            "hbm_gbs_step": ALG_BYTES_PER_ELEM.get(integrator, 0) * n_shard * pipe["attempts"] / (pipe["ms"] * 1e-3) / 1e9,
                            "roofline": stage_roofline(pipe)}
        if head is pipe:
            roofline = stage_roofline(pipe)
            path = "stage/RHS/finish pipeline"
        elif rhs_kind != "diag":
            roofline = stage_roofline(head)
            roofline["kernel"] = "stage_l96_kernel<M> (stage accumulate fused with the Lorenz-96 stencil through a shared-memory tile; all %d launches)" % head["prof"]["stage"]["launches"]
            roofline["traffic"] = traffic_db.get("stage_l96_kernel_%s_2p%d" % (integrator, lg))
            path = "stage+stencil fused (built-in Lorenz-96), finish kernel"
        else:
            devloop = ctx.get("device_loop") != 0 and (world == 1 or bool(ctx.get("p2p")))
            fu = prof["fused"]
            a = gbs(fu)
            attempts_instr = fu["bytes"] / (8.0 * n_shard * 5) if n_shard else 0.0  # 5 vector passes per attempt
            if devloop:
                path = "fused_run (element-local built-in RHS: all K steps in one persistent cooperative kernel, controller on the device)"
                kname = ("fused_run_kernel<PAT,RHS> (persistent cooperative kernel: every attempt = all stages, RHS, yNew, error norm, "
                         "grid barrier, device-side controller; one launch runs the K steps)")
            else:
                path = "fused_attempt (element-local built-in RHS)"
                kname = "fused_attempt_kernel<PAT,RHS> (whole attempt of an element-local IVP in one kernel: all stages, RHS, yNew, error norm)"
            roofline = {"bound": "hbm", "kernel": kname,
                        "achieved": a, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": a / peak, "frac_of_nominal_8000": a / 8000.0,
                        "traffic": (traffic_db.get("fused_run_kernel_per_attempt_%s_2p%d" % (integrator, lg), 0.0) * attempts_instr / max(1, fu["launches"]) or None)
                        if devloop else traffic_db.get("fused_attempt_kernel_%s_2p%d" % (integrator, lg)),
                        "launches": fu["launches"], "attempts_in_launches": attempts_instr, "us_per_attempt": 1e3 * fu["ms"] / max(1.0, attempts_instr),
                        "avg_launch_us": 1e3 * fu["ms"] / max(1, fu["launches"]), "algorithmic_bytes_per_launch": fu["bytes"] / max(1, fu["launches"]),
                        "instrumented_ms_per_step": prof["instrumented_ms"] / args.steps,
                        "kernel_time_share_of_step": fu["ms"] / prof["instrumented_ms"] if prof["instrumented_ms"] > 0 else None}
        l96_obj = None
        if l96_attempt is not None:
            if "error" in l96_attempt:
                l96_obj = l96_attempt
            else:
                fu = l96_attempt["prof"]["fused"]
                l96_obj = {"note": "experimental knob fuse_stencil_attempt=1: the whole attempt (all stages, Lorenz-96 stencil, yNew, error norm) in one "
                                   "kernel over overlapped tiles (csrc/stencil_attempt.cuh); algorithmic bytes = 4 vector passes per attempt",
                           "value": args.steps * world / (l96_attempt["ms_max"] * 1e-3), "ms_per_step": l96_attempt["ms_max"] / args.steps,
                           "attempts": l96_attempt["attempts"], "rejected": l96_attempt["rejected"], "gpu_launches": l96_attempt["launches"],
                           "t_reached": l96_attempt["t"], "hbm_gbs": gbs(fu), "frac_of_peak": gbs(fu) / peak, "launches": fu["launches"],
                           "avg_launch_us": 1e3 * fu["ms"] / max(1, fu["launches"])}
        line = {
            "metric": "rk_steps_per_sec", "value": args.steps * world / (ms_max * 1e-3), "unit": "RK steps/s (x 2^%d-element shard, summed over GPUs)" % lg,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "integrator": integrator, "rhs": rhs_kind, "elems_per_gpu": n_shard, "elems_global": n_global,
                       "options": OPTS, "l2": "working set (>= 10 vectors x %d MiB per GPU) exceeds the 126 MB L2; no flush" % (n_shard * 8 >> 20),
                       "sharding": "contiguous shards, 1 all-reduce(sum, 1 x f64) of the error norm per attempt" if world > 1 else "single GPU, no collective",
                       "vec_width": ctx.get("vec_width"), "ctas_per_sm": ctx.get("ctas_per_sm"), "finish_ctas_per_sm": ctx.get("finish_ctas_per_sm"),
                       "fuse_pointwise": ctx.get("fuse_pointwise"), "fused_ctas_per_sm": ctx.get("fused_ctas_per_sm"),
                       "spin_readback": ctx.get("spin_readback"), "l2_hints": ctx.get("l2_hints"), "device_loop": ctx.get("device_loop"),
                       "error_norm_allreduce": ("in-kernel peer mailboxes over NVLink (CUDA IPC)" if ctx.get("p2p") else "ncclAllReduce") if world > 1 else None},
            "attempts": attempts, "attempts_per_sec": attempts * world / (ms_max * 1e-3), "rejected": head["rejected"],
            "t_reached": t_now, "dt_next": dt_next,
            "gpu_launches": launches, "collectives": head["collectives"],
            "path": path,
            "roofline": roofline,
            "pipeline": pipeline_obj,
            "jit_rhs": jit_obj,
            "l96_attempt": l96_obj,
            "trajectory_consumers": quad_obj,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_steps * world / (e2e_ms_max * 1e-3), "unit": "RK steps/s (solveODE on host buffers: H2D y0+lambda, solve, D2H states)",
                    "h2d_bytes_per_step": h2d / max(1, e2e_steps), "d2h_bytes_per_step": d2h / max(1, e2e_steps), "steps_per_solve": e2e_steps / reps,
                    "ms_per_solve": e2e_ms_max / reps},
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ======================================================================================================
# CPU: oracle port (the reference cannot be built: Nim, no toolchain) — bench.py's only use of oracle/
# ======================================================================================================
def cpu_baseline_sample(integrator: str, n: int, steps: int):
    import oracle as O

    lam, y0 = problem_arrays(n, 0, n)
    t0 = time.time()
    s = O.solve_vector(integrator, O.rhs_diag_linear(lam), y0, [0.0, 1e12], O.new_options(**OPTS), max_steps=steps)
    wall = time.time() - t0
    out = {"value": s.stats.steps / s.stats.seconds, "unit": "RK steps/s at 2^%d elements" % int(np.log2(n)), "cores": 1, "kind": "port",
           "sample": "%d accepted %s steps (+2 start-up RHS evaluations) at N=2^%d on 1 of %d host cores; %.1f s" % (
               s.stats.steps, integrator, int(np.log2(n)), os.cpu_count() or 0, wall)}
    try:  # courtesy row: NOT the reference's behaviour — the attempt fused into one pass, all host cores
        _, fs = O.fused_mt_solve_diag(integrator, lam, y0, 1e12, O.new_options(**OPTS), max_steps=10)
        out["courtesy_fused_multithreaded"] = {"value": fs.steps / fs.seconds, "unit": out["unit"], "cores": int(fs.threads),
                                               "note": "same IVP, whole attempt fused into one pass per element (5 vector passes), std::thread over all host cores, scalar -O2 code; "
                                                       "element-wise results bit-identical to the port. numericalnim itself is single-threaded and allocates a Vector per operator."}
    except Exception as e:  # noqa: BLE001
        out["courtesy_fused_multithreaded"] = {"error": str(e)[:200]}
    return out


def run_reference(args):
    """The reference's own CPU implementation of the path, as faithfully as it can be had here: the C++
    oracle port (allocating one-pass-per-operator Vectors, sequential sum, -O2, no FMA). Single thread because
    numericalnim's ode.nim / utils.nim have no threading construct — that is every thread the reference uses."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import oracle as O

    integrator, rhs_kind, lg = WORKLOADS[args.workload]
    if rhs_kind != "diag":
        raise SystemExit("reference arm implements the diag-linear workloads")
    n_full = 1 << (args.log2n or lg)
    # calibrate on a small sample, then pick the largest power-of-two N_s whose K+W steps fit the budget
    lam, y0 = problem_arrays(1 << 18, 0, 1 << 18)
    c = O.solve_vector(integrator, O.rhs_diag_linear(lam), y0, [0.0, 1e12], O.new_options(**OPTS), max_steps=2)
    per_elem_step = c.stats.seconds / 2 / (1 << 18) * 3.0  # large vectors are ~3x slower per element (page faults)
    budget = args.cpu_budget_s
    n_s = n_full
    while n_s > (1 << 16) and per_elem_step * n_s * (args.steps + args.warmup) > budget:
        n_s >>= 1
    lam, y0 = problem_arrays(n_s, 0, n_s)
    O.solve_vector(integrator, O.rhs_diag_linear(lam), y0, [0.0, 1e12], O.new_options(**OPTS), max_steps=max(1, args.warmup))
    t0 = time.time()
    s = O.solve_vector(integrator, O.rhs_diag_linear(lam), y0, [0.0, 1e12], O.new_options(**OPTS), max_steps=args.steps)
    wall = time.time() - t0
    scale = n_full / n_s  # time assumed linear in N (bandwidth-bound) — favours the CPU, see DESIGN.md §6
    secs = s.stats.seconds * scale
    value = s.stats.steps / secs
    sample = "%d accepted %s steps at N_s=2^%d (%s of the 2^%d workload, time scaled x%g), 1 of %d host cores, %.1f s wall" % (
        s.stats.steps, integrator, int(np.log2(n_s)), "all" if n_s == n_full else "1/%d" % int(scale), int(np.log2(n_full)), scale, os.cpu_count() or 0, wall)
    line = {"impl": "reference", "metric": "rk_steps_per_sec", "value": value,
            "unit": "RK steps/s (x 2^%d-element shard, summed over GPUs)" % int(np.log2(n_full)), "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, s.stats.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "integrator": integrator, "rhs": rhs_kind, "elems_per_gpu": n_full, "options": OPTS,
                       "note": "CPU arm: one shard on one host thread (the reference is single-threaded); not multiplied by n_gpus"},
            "cpu_baseline": {"value": value, "unit": "RK steps/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "RK steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    ap.add_argument("--l96-attempt", action="store_true", help="Lorenz-96 workloads: also time the experimental one-kernel attempt (knob fuse_stencil_attempt)")
    ap.add_argument("--no-fuse", action="store_true", help="headline = the stage/RHS/finish pipeline even for element-local built-in RHS")
    ap.add_argument("--cpu-budget-s", type=float, default=120.0)
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
