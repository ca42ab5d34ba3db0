"""Second, independent restatement of numericalnim's explicit-RK path in pure Python (IEEE doubles, no
numpy in the arithmetic) — TEST INFRASTRUCTURE used to cross-check the C++ oracle bit for bit and to
generate the golden fixtures in tests/golden/ (see tests/golden/make_golden.py).

It differs on purpose from oracle/rk_oracle.hpp in structure: the Butcher coefficients are not typed in
here at all — they are read from ``tests/golden/tableaux.json``, which make_golden.py extracted
mechanically from the reference source (``ode.nim:241-282, 311-352, 381-443``) in the build container.

Reference lines restated: ode.nim:57-76 (retry loop), :180-189 (RK4), :237-468 (pairs), :471-586 (driver);
utils.nim:59-64,113-118,171-197,214-250 (Vector ops), :273-279 (hermiteSpline), :498-507 (linspace).
"""
from __future__ import annotations

import json
import math
import os

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Vec:
    """Vector[float] (utils.nim:14-17) with the operators the RK loop uses."""

    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [float(x) for x in c]

    def __len__(self):
        return len(self.c)

    def _chk(self, o):
        if len(self.c) != len(o.c):
            raise ValueError("Vectors must have the same size.")

    def __add__(self, o):
        self._chk(o)
        return Vec([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        self._chk(o)
        return Vec([a - b for a, b in zip(self.c, o.c)])

    def __rmul__(self, d):
        return Vec([a * d for a in self.c])

    def __mul__(self, d):
        return Vec([a * d for a in self.c])

    def __neg__(self):
        return Vec([-a for a in self.c])


def vabs(v):
    return Vec([abs(a) for a in v.c]) if isinstance(v, Vec) else abs(v)


def dot_add(d, v):  # `+.`
    return Vec([a + d for a in v.c]) if isinstance(v, Vec) else d + v


def dot_mul(a, b):  # `*.`
    return Vec([x * y for x, y in zip(a.c, b.c)]) if isinstance(a, Vec) else a * b


def dot_div(a, b):  # `/.`
    if isinstance(a, Vec):
        out = []
        for x, y in zip(a.c, b.c):
            try:
                out.append(x / y)
            except ZeroDivisionError:
                out.append(math.nan if x == 0 or x != x else math.copysign(math.inf, x) * math.copysign(1.0, y))
        return Vec(out)
    return a / b


def size(v):
    return len(v.c) if isinstance(v, Vec) else 1


def vsum(v):
    if not isinstance(v, Vec):
        return v
    r = 0.0
    for a in v.c:
        r = r + a
    return r


def nmin(x, y):
    return x if x <= y else y


def nmax(x, y):
    return x if y <= x else y


def linspace(x1, x2, n):
    dx = (x2 - x1) / float(n - 1)
    r = [x1]
    for i in range(1, n - 1):
        r.append(x1 + dx * float(i))
    r.append(x2)
    return r


def hermite(x, x1, x2, y1, y2, dy1, dy2):
    t = (x - x1) / (x2 - x1)
    h00 = (1.0 + 2.0 * t) * ((1.0 - t) * (1.0 - t))
    h10 = t * ((1.0 - t) * (1.0 - t))
    h01 = (t * t) * (3.0 - 2.0 * t)
    h11 = t * (t * t) - t * t
    return h00 * y1 + h10 * (x2 - x1) * dy1 + h01 * y2 + h11 * (x2 - x1) * dy2


def options(dt=1e-4, absTol=1e-4, relTol=1e-4, dtMax=1e-2, dtMin=1e-4, scaleMax=4.0, scaleMin=0.1, tStart=0.0):
    if abs(dtMax) < abs(dtMin):
        raise ValueError("dtMin must be less than dtMax")
    if abs(scaleMax) < 1:
        raise ValueError("scaleMax must be bigger than 1")
    if 1 < abs(scaleMin):
        raise ValueError("scaleMin must be smaller than 1")
    return dict(dt=abs(dt), absTol=abs(absTol), relTol=abs(relTol), dtMax=abs(dtMax), dtMin=abs(dtMin),
                scaleMax=abs(scaleMax), scaleMin=abs(scaleMin), tStart=tStart)


_TAB = None


def tableaux():
    global _TAB
    if _TAB is None:
        with open(os.path.join(_GOLDEN, "tableaux.json")) as fh:
            raw = json.load(fh)
        _TAB = {m: {k: float.fromhex(v) for k, v in d.items()} for m, d in raw.items()}
    return _TAB


def rk4_step(f, t, y, fsal, dt, o, st):
    k1 = f(t, y)
    k2 = f(t + 0.5 * dt, y + 0.5 * dt * k1)
    k3 = f(t + 0.5 * dt, y + 0.5 * dt * k2)
    k4 = f(t + dt, y + dt * k3)
    yn = y + dt / 6.0 * (k1 + 2.0 * (k2 + k3) + k4)
    return yn, yn, dt, 0.0


def _pair_step(name, nst, order, direct):
    def step(f, t, y, fsal, dt, o, st):
        T = tableaux()[name]
        lim = 0
        while lim < 2:
            k = [None, fsal]  # 1-based
            for s in range(2, nst + 1):
                acc = T[f"a{s}1"] * k[1]
                for j in range(2, s):
                    acc = acc + T[f"a{s}{j}"] * k[j]
                k.append(f(t + dt * T[f"c{s}"], y + dt * acc))
            nb = nst - 1
            acc = T["b1"] * k[1]
            for j in range(2, nb + 1):
                acc = acc + T[f"b{j}"] * k[j]
            yn = y + dt * acc
            acc = T["bHat1"] * k[1]
            for j in range(2, nst + 1):
                acc = acc + T[f"bHat{j}"] * k[j]
            if direct:
                ey = dt * acc
            else:
                ey = yn - (y + dt * acc)
            tol = dot_add(o["absTol"], o["relTol"] * vabs(yn))
            e1 = dot_div(ey, tol)
            sq = dot_mul(e1, e1)
            err = math.sqrt(1 / float(size(e1)) * vsum(sq))
            st["attempts"] += 1
            if err <= 1:
                break
            st["rejected"] += 1
            dt = dt * nmin(4, nmax(0.125, 0.9 * math.pow(1 / err, 1 / order)))
            if abs(dt) < o["dtMin"]:
                dt = o["dtMin"]
                lim += 1
                st["limiter_hits"] += 1
            elif o["dtMax"] < abs(dt):
                dt = o["dtMax"]
        return yn, k[nst], dt, err

    return step


METHODS = {
    "rk4": (rk4_step, False, 4.0, False),
    "dopri54": (_pair_step("dopri54", 7, 5, False), True, 5.0, True),
    "tsit54": (_pair_step("tsit54", 7, 5, True), True, 5.0, True),
    "vern65": (_pair_step("vern65", 9, 6, False), True, 6.0, True),
}


def solve(f, y0, tspan, o=None, integrator="dopri54"):
    """ODESolver (ode.nim:471-586). Returns (t list, y list, stats dict with 'trace')."""
    o = o or options()
    step, use_fsal, order, adaptive = METHODS[integrator.lower()]
    tspan = sorted(tspan)
    st = dict(attempts=0, rejected=0, limiter_hits=0, steps=0, trace=[])
    t0 = o["tStart"]
    tpos = [x for x in tspan if x > t0]
    tneg = [x for x in tspan if x < t0][::-1]
    ypos, yneg, yzero, tzero = [], [], [], []
    y = y0
    if t0 in tspan:
        yzero.append(y)
        tzero.append(t0)
    dt_init = math.sqrt(o["dtMax"] * o["dtMin"]) if adaptive else o["dt"]
    dense = len(tspan) != 2

    def run(fn, t, y, tend, targets, sign, out):
        fsal = fn(t, y)
        last = (t, y, fn(t, y) if sign > 0 else fsal)
        dt = dt_init
        di = 0
        err = 0.0
        while t < tend:
            if dense:
                if len(targets) - 1 < di:
                    break
                while sign * targets[di] <= t:
                    out.append(hermite(sign * targets[di], last[0], t, last[1], y, last[2], fsal if use_fsal else fn(t, y)))
                    di += 1
                    if len(targets) - 1 < di:
                        break
            dt = nmin(dt, tend - t)
            if dense:
                last = (t, y, fsal if use_fsal else fn(t, y))
            a0 = st["attempts"]
            tb = t
            y, fsal, dt, err = step(fn, t, y, fsal, dt, o, st)
            t += dt
            st["steps"] += 1
            st["trace"].append((sign * tb, dt, err, st["attempts"] - a0))
            if adaptive:
                if err == 0.0:
                    dt *= 5
                else:
                    dt = dt * nmin(4, nmax(0.125, 0.9 * math.pow(1 / err, 1 / order)))
                if dt < o["dtMin"]:
                    dt = o["dtMin"]
                elif o["dtMax"] < dt:
                    dt = o["dtMax"]
        out.append(y)

    if tpos:
        run(f, t0, y, max(tpos), tpos, 1.0, ypos)
    if tneg:
        run(lambda t, yy: -f(-t, yy), -t0, y0, -min(tneg), tneg, -1.0, yneg)
    return tneg[::-1] + tzero + tpos, yneg[::-1] + yzero + ypos, st


def rhs_scale(c):
    return lambda t, y: c * y


def rhs_diag(lam):
    L = Vec(lam)
    return lambda t, y: -dot_mul(L, y)


def rhs_l96(F):
    def f(t, y):
        n = len(y.c)
        c = y.c
        return Vec([((c[(i + 1) % n] - c[(i - 2) % n]) * c[(i - 1) % n] - c[i]) + F for i in range(n)])

    return f
