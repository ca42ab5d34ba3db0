#!/usr/bin/env python
"""Per-kernel roofline model from the shipped SASS (no GPU needed): fp64 instructions per element (static count of the
W-wide main body + scalar tail, divided by W + 1 elements) against the HBM time of the kernel's algorithmic bytes.
B200: 148 SMs x 64 fp64 lanes x 1.965 GHz = 1.86e13 DMUL/DADD/DFMA per second; HBM = MEASURED_PEAKS.json.
Prints a markdown table; profiles/r01_kernel_model.md keeps the round-1 copy next to the measured times."""
import collections
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "numericalnim_b200", "lib", "obj")
FP64_RATE = 148 * 64 * 1.965e9


def kernels():
    out = {}
    for path in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
        if path.endswith("kernels_src.o"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        fn = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = m.group(1)
                out[fn] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and fn:
                out[fn][m.group(1)] += 1
    return out


def registers():
    res = ""
    for path in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
        if path.endswith("kernels_src.o"):
            continue
        res += subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True, check=True).stdout
    return {m[0]: int(m[1]) for m in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res)}


def fp64(c):
    return sum(v for k, v in c.items() if k.startswith(("DADD", "DMUL", "DFMA", "DSETP", "MUFU")))


# (label, mangled-name fragment, W, U, vector passes per element, elements, measured us or None, source of the measurement)
ROWS = [
    ("stage_kernel m=5 (DOPRI54 stage 6)", "stage_kernelILi5ELi4ELi1ELb0ELi256ELi1E", 4, 1, 7, 1 << 23, 59.6, "r01_launches_cfg2_with_jit_rhs.csv"),
    ("finish_kernel DOPRI54 (6 k-streams)", "finish_kernelILi6ELi4ELi1ELb0ELi0ELi256E", 4, 1, 7, 1 << 23, 88.2, "r01_bench_cfg2_n1_jit_rhs.json (in-step)"),
    ("fused_attempt DOPRI54 diag", "fused_attempt_kernelILi0ELi1ELi4ELi256ELi0E", 4, 1, 5, 1 << 23, 64.0, "r01_bench_fused.json"),
    # the device-resident loop runs the same per-element code as fused_attempt (its own static count also holds the
    # controller: pow, sqrt, the partial-sum loops), so the per-element figure is taken from the attempt kernel
    ("fused_run DOPRI54 diag (per attempt, 2^23)", "fused_attempt_kernelILi0ELi1ELi4ELi256ELi0E", 4, 1, 5, 1 << 23, 54.6, "r01_bench_cfg2_n1_jit_rhs.json"),
    ("fused_attempt Tsit54 diag", "fused_attempt_kernelILi2ELi1ELi4ELi256ELi0E", 4, 1, 5, 1 << 23, None, ""),
    ("fused_run Vern65 diag (per attempt, 2^24)", "fused_attempt_kernelILi3ELi1ELi4ELi256ELi0E", 4, 1, 5, 1 << 24, 156.0, "r01_bench_cfg4_1gpu_device_loop.json"),
    ("cumtrapz_kernel (per time point)", "cumtrapz_kernelILi4ELi4ELi256E", 4, 4, 2, 1 << 23, 668.1 / 33, "r01_quadrature_2p23_m33.json"),
    ("simpson_scan_kernel (per pair)", "simpson_scan_kernelILi4ELi256E", 4, 1, 3, 1 << 23, 485.0 / 16, "r01_ncu_full_quadrature_2p23.csv"),
    ("hermite_many_kernel (per sample)", "hermite_many_kernelILi4ELi256E", 4, 1, 2, 1 << 23, 1311.0 / 66, "r01_quadrature_2p23_m33.json"),
    # experimental one-kernel Lorenz-96 attempt (stencil_attempt.cuh): 4 elements per thread, no scalar tail (W*U + 1 = 4
    # with W = 3 below is only how this table divides); 2 % of the tile is redundant overlap. Not yet run on a GPU.
    ("l96_attempt Tsit54 (whole attempt, 2^24; the default stage+stencil path takes 841 us)", "l96_attempt_kernelILi2ELi2ELi256ELb0E", 3, 1, 4, int((1 << 24) * 1024 / 1004), None, ""),
    ("l96_attempt Vern65 (whole attempt, 2^24)", "l96_attempt_kernelILi3ELi2ELi256ELb0E", 3, 1, 4, int((1 << 24) * 1024 / 1000), None, ""),
]


def main():
    peak = 6460.2e9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"]) * 1e9
    ks = kernels()
    regs = registers()
    print("| kernel | registers -> resident CTAs/SM (256 threads) | fp64-pipe instr / element | fp64-pipe time | HBM time (algorithmic bytes) | bound | measured | measured / max(model) |")
    print("|---|---|---|---|---|---|---|---|")
    for label, frag, w, u, passes, n, meas, src in ROWS:
        match = [c for fn, c in ks.items() if frag in fn]
        if not match:
            print(f"| {label} | (not found: {frag}) | | | | | | |")
            continue
        c = match[0]
        r = [v for fn, v in regs.items() if frag in fn][0]
        occ = f"{r} -> {min(8, 65536 // (max(r, 1) * 256))}"
        per_elem = fp64(c) / float(w * u + 1)   # static body = W*U elements + the scalar tail's 1
        t_fp = per_elem * n / FP64_RATE * 1e6
        t_hbm = passes * 8.0 * n / peak * 1e6
        bound = "fp64 issue" if t_fp > t_hbm else "HBM"
        ms = f"{meas:.1f} us ({src})" if meas else "—"
        ratio = f"{meas / max(t_fp, t_hbm):.2f}" if meas else "—"
        print(f"| {label} | {occ} | {per_elem:.0f} | {t_fp:.1f} us | {t_hbm:.1f} us | {bound} | {ms} | {ratio} |")


if __name__ == "__main__":
    sys.exit(main())
