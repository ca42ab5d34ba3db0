#!/usr/bin/env python
"""Generate the committed golden fixtures (run ONCE in the build container, where /root/reference is
mounted; the GPU box has no /root/reference, so tests only ever read the JSON this script wrote).

  tableaux.json       Butcher coefficients of DOPRI54 / Tsit54 / Vern65, extracted mechanically from the
                      `const` blocks of /root/reference/src/numericalnim/ode.nim (lines 240-282, 310-352,
                      380-443) and stored as C99 hex floats — no hand transcription.
  trajectories.json   Trajectories / step sequences produced by tests/pyref.py (the independent pure-Python
                      restatement driven by tableaux.json) for the reference's own test IVP
                      (tests/test_ode.nim:5-16) and for the controller-exercising cases of SURVEY.md
                      Appendix B. Stored as hex floats so comparisons are bit-exact.

The reference itself cannot be executed (no Nim toolchain), so these are "restatement-generated" vectors:
they pin the oracle and the CUDA path to each other and to the reference's literal constants; the
reference's own pass/fail criteria (t == tspan, |y - exp(-0.1 t)| <= tol) are asserted separately.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference/src/numericalnim/ode.nim"


def extract_tableaux():
    src = open(REF).read()
    out = {}
    for name, proc in (("dopri54", "DOPRI54_step"), ("tsit54", "TSIT54_step"), ("vern65", "VERN65_step")):
        start = src.index(f"proc {proc}")
        block = src[start:src.index("let absTol", start)]
        block = block[block.index("const"):]
        ns = {}
        for line in block.splitlines():
            m = re.match(r"\s+([A-Za-z]+\d+)\s*=\s*([^#]+?)\s*(#.*)?$", line)
            if not m:
                continue
            key, expr = m.group(1), m.group(2)
            if not re.fullmatch(r"[-0-9.eE/ A-Za-z]+", expr):
                raise SystemExit(f"unexpected expression {expr!r}")
            ns[key] = float(eval(expr, {"__builtins__": {}}, dict(ns)))
        out[name] = {k: v.hex() for k, v in ns.items()}
    return out


def hexlist(xs):
    return [float(x).hex() for x in xs]


def main():
    tabs = extract_tableaux()
    with open(os.path.join(HERE, "tableaux.json"), "w") as fh:
        json.dump(tabs, fh, indent=1, sort_keys=True)
    import pyref as R

    R._TAB = None
    cases = {}

    def add(name, integrator, rhs_desc, f, y0, tspan, opt_kwargs, vector):
        o = R.options(**opt_kwargs)
        t, y, st = R.solve(f, R.Vec(y0) if vector else y0[0], tspan, o, integrator)
        cases[name] = dict(
            integrator=integrator, rhs=rhs_desc, y0=hexlist(y0), tspan=hexlist(tspan), options=opt_kwargs, vector=vector,
            t=hexlist(t), y=[hexlist(v.c if vector else [v]) for v in y],
            steps=st["steps"], attempts=st["attempts"], rejected=st["rejected"], limiter_hits=st["limiter_hits"],
            trace_dt=hexlist(r[1] for r in st["trace"][:64]), trace_err=hexlist(r[2] for r in st["trace"][:64]),
            trace_attempts=[r[3] for r in st["trace"][:64]],
        )

    ts100 = R.linspace(-10.0, 10.0, 100)
    # tests/test_ode.nim Vector cases (y0 = [1,1,1], f = -0.1*y), default options and ooVector.
    for m in ("dopri54", "tsit54", "vern65", "rk4"):
        add(f"testode_vec3_{m}_default", m, {"kind": "scale", "c": -0.1}, R.rhs_scale(-0.1), [1.0, 1.0, 1.0], ts100, {}, True)
        add(f"testode_vec3_{m}_oovector", m, {"kind": "scale", "c": -0.1}, R.rhs_scale(-0.1), [1.0, 1.0, 1.0], ts100,
            dict(relTol=1e-8, dt=1e-2), True)
    # scalar cases, adaptive only (rk4 default = 200k python steps: covered by the C++ oracle test instead)
    for m in ("dopri54", "tsit54", "vern65"):
        add(f"testode_scalar_{m}_default", m, {"kind": "scale", "c": -0.1}, R.rhs_scale(-0.1), [1.0], ts100, {}, False)
    # controller really moves
    lam8 = [0.1 + 9.9 * i / 7 for i in range(8)]
    optB = dict(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8)
    for m in ("dopri54", "tsit54", "vern65"):
        add(f"diag8_{m}", m, {"kind": "diag", "lam": hexlist(lam8)}, R.rhs_diag(lam8), [1.0] * 8, [0.0, 2.0], optB, True)
    y0 = [8.0] * 40
    y0[0] = 8.01
    for m in ("dopri54", "tsit54", "vern65"):
        add(f"l96_40_{m}", m, {"kind": "l96", "F": 8.0}, R.rhs_l96(8.0), y0, [0.0, 1.0], optB, True)
    # limiter path (ode.nim:72-74)
    optL = dict(absTol=1e-8, relTol=1e-8, dtMax=0.1, dtMin=1e-3)
    for m in ("dopri54", "tsit54", "vern65"):
        add(f"limiter4_{m}", m, {"kind": "diag", "lam": hexlist([1, 10, 100, 1000])}, R.rhs_diag([1.0, 10.0, 100.0, 1000.0]),
            [1.0] * 4, [0.0, 0.05], optL, True)
    # dense output + backward time with a moving controller
    tsd = R.linspace(-1.0, 1.5, 11)
    for m in ("dopri54", "tsit54", "vern65", "rk4"):
        add(f"dense_diag8_{m}", m, {"kind": "diag", "lam": hexlist(lam8)}, R.rhs_diag(lam8), [1.0 + 0.1 * i for i in range(8)], tsd,
            dict(absTol=1e-6, relTol=1e-6, dtMax=0.5, dtMin=1e-8, dt=1e-2), True)
    # edge cases of the driver's tspan handling (ode.nim:476-487, 499-502, 542, 585-586)
    optE = dict(absTol=1e-6, relTol=1e-6, dtMax=0.25, dtMin=1e-8, dt=1e-2)
    y0e = [1.0 + 0.1 * i for i in range(8)]
    edge = {
        "only_tstart": ([0.0], {}),                         # nothing to integrate: returns (tStart, y0)
        "tstart_inside_not_listed": ([0.0, 0.5, 1.0], dict(tStart=0.25)),   # both directions, tZero empty, dense
        "duplicates": ([0.0, 1.0, 1.0, 2.0], {}),           # dense loop emits the duplicate, final add(y) adds one more
        "negative_only_len2": ([-1.0, -0.5], {}),           # len 2 => no dense: 2 times, 1 state
        "straddle_len2": ([-0.5, 0.5], {}),                 # len 2, tStart between: one state per direction
        "unsorted": ([1.0, -1.0, 0.0, 0.5], {}),            # sorted by solveODE (ode.nim:609)
        "nonzero_tstart_listed": ([1.0, 1.5, 2.0], dict(tStart=1.0)),
    }
    for ename, (ts, extra) in edge.items():
        for m in ("tsit54", "rk4"):
            add(f"edge_{ename}_{m}", m, {"kind": "diag", "lam": hexlist(lam8)}, R.rhs_diag(lam8), y0e, ts, dict(optE, **extra), True)
    with open(os.path.join(HERE, "trajectories.json"), "w") as fh:
        json.dump(cases, fh, indent=0, sort_keys=True)
    print("wrote", len(tabs), "tableaux and", len(cases), "trajectories")


if __name__ == "__main__":
    main()
