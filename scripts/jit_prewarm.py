#!/usr/bin/env python
"""Compile, on the CPU box, every run-time compiled right-hand side the GPU tests and bench.py use, into the
in-tree cache (.jitcache/, shipped to the GPU box with the snapshot). NVRTC needs no GPU; the cache key holds the
NVRTC version and the contents of kernels.cuh, so a stale entry is never used. Purely a time saver: a miss on the
GPU box just compiles there (~3 s per translation unit)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200RK_JIT_CACHE", os.path.join(ROOT, ".jitcache"))

import numericalnim_b200 as nn  # noqa: E402

ALL = (-1, 0, 1, 2, 3, 4)
UNITS = [  # (expression, n_vec, n_scalar, patterns)
    ("c0*y*(1.0 - y/p0) + c1*t", 1, 2, ALL),      # tests/test_gpu_rhs_from_source.py LOGISTIC
    ("-(p0*y) + p1*(c0*t)", 2, 1, ALL),           # FORCED
    ("c0*y", 0, 1, (-1, 0)),                      # SCALE; examples/cpp_host_demo.cpp
    ("-(p0*y)", 1, 0, (-1, 0, 3)),                # bench.py jit leg (dopri54, vern65), builtin-equality test
    ("c0*y + c1", 0, 2, (-1,)),
    ("p0*y", 1, 0, (-1,)),
    ("-y*exp(-c0*t) + sin(p0)", 1, 1, (-1,)),
]

STENCIL_UNITS = [  # (expression, radius_left, radius_right, n_vec, n_scalar, patterns) — tests/test_gpu_stencil_from_source.py, bench.py
    ("((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, 0, 1, ALL),
    ("c0*((Y(-1) - 2.0*Y(0)) + Y(1))*p0 + c1*t", 1, 1, 1, 2, (-1, 0, 2, 3)),
    ("-(c0*(Y(0) - Y(-1)))", 1, 0, 0, 1, (-1, 0, 2, 3)),
    ("((Y(-3) + Y(2)) - 2.0*Y(0))*c0 - Y(0)*Y(0)*Y(0)*c1", 3, 2, 0, 2, (-1, 0, 2, 3)),
    ("Y(-2) + Y(1)", 2, 1, 0, 0, (-1,)),
    ("c0*(Y(3) - Y(0))", 0, 3, 0, 1, (-1, 0, 2, 3)),
    ("((Y(-8) + Y(8)) - 2.0*Y(0))*c0", 8, 8, 0, 1, (-1, 0, 2, 3)),
    ("c0*Y(0) + c1*t", 0, 0, 0, 2, (-1, 0, 2, 3)),
    ("c0*((Y(-1) - 2.0*Y(0)) + Y(1))", 1, 1, 0, 1, (-1, 2)),          # examples/c_extras_demo.c
]

if __name__ == "__main__":
    t0 = time.time()
    n = 0
    for expr, nv, nc, pats in UNITS:
        for p in pats:
            nn.jitCompileOnly(expr, nv, nc, p)
            n += 1
    for expr, rl, rr, nv, nc, pats in STENCIL_UNITS:
        for p in pats:
            nn.jitStencilCompileOnly(expr, rl, rr, nv, nc, p)
            n += 1
    print(f"{n} units in {time.time() - t0:.1f} s -> {os.environ['B200RK_JIT_CACHE']}")
