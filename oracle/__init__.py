"""ctypes front-end of the CPU oracle (``oracle/rk_oracle.hpp``).

TEST INFRASTRUCTURE ONLY. Importable from ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``. The product package
``numericalnim_b200`` must never import this module.

The oracle restates numericalnim's ``ode.nim`` / ``utils.nim`` arithmetic on the CPU (see the header of
``rk_oracle.hpp`` for the pinning status: pinned on the reference's own test cases, PARITY UNPINNED on the
step-rejection / dtMin-limiter path).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

RHS_SCALE, RHS_DIAG_LINEAR, RHS_LORENZ96, RHS_CALLBACK = 0, 1, 2, 3


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ -O2 -ffp-contract=off)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("rk_oracle.hpp", "quad_oracle.hpp", "oracle_capi.cpp", "Makefile")
    )
    if force or src_newer:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


class OracleOptions(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("dt", "dtMax", "dtMin", "tStart", "absTol", "relTol", "scaleMax", "scaleMin")]


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_long) for n in ("steps", "attempts", "rejected", "limiter_hits", "rhs_evals", "nan_guard")] + [
        ("seconds", C.c_double)
    ]


class StepRecord(C.Structure):
    _fields_ = [("t", C.c_double), ("dt_used", C.c_double), ("error", C.c_double), ("attempts", C.c_int), ("_pad", C.c_int)]


RHS_CB = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t, C.c_void_p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_vector_sum.restype = C.c_double
        L.oracle_vector_norm.restype = C.c_double
        L.oracle_vector_dot.restype = C.c_double
        L.oracle_vector_sum.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_vector_norm.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.oracle_vector_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_is_close_scalar.argtypes = [C.c_double, C.c_double, C.c_double]
        L.oracle_is_close_vec.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]
        _lib = L
    return _lib


class ValueError_(ValueError):
    """Stands for Nim's ValueError raised by the reference."""


def _check(rc: int):
    if rc != 0:
        raise ValueError_(lib().oracle_last_error().decode())


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def new_options(dt=1e-4, absTol=1e-4, relTol=1e-4, dtMax=1e-2, dtMin=1e-4, scaleMax=4.0, scaleMin=0.1, tStart=0.0) -> OracleOptions:
    o = OracleOptions()
    _check(lib().oracle_new_options(C.byref(o), *(C.c_double(float(x)) for x in (dt, absTol, relTol, dtMax, dtMin, scaleMax, scaleMin, tStart))))
    return o


def linspace(x1: float, x2: float, n: int) -> np.ndarray:
    out = np.empty(max(n, 1), dtype=np.float64)
    _check(lib().oracle_linspace(C.c_double(x1), C.c_double(x2), C.c_long(n), _p(out)))
    return out


@dataclass
class Rhs:
    kind: int
    param: np.ndarray | None = None
    scalar: float = 0.0
    callback: object = None  # python callable (t, y: ndarray) -> ndarray

    def _cb(self):
        if self.kind != RHS_CALLBACK:
            return C.cast(None, RHS_CB), None
        fn = self.callback

        def tramp(t, yp, op, n, _user):
            y = np.ctypeslib.as_array(yp, shape=(n,))
            out = np.ctypeslib.as_array(op, shape=(n,))
            out[:] = fn(t, y.copy())

        cb = RHS_CB(tramp)
        return cb, cb


def rhs_scale(c: float) -> Rhs:
    return Rhs(RHS_SCALE, None, float(c))


def rhs_diag_linear(lam) -> Rhs:
    return Rhs(RHS_DIAG_LINEAR, _f64(lam))


def rhs_lorenz96(F: float = 8.0) -> Rhs:
    return Rhs(RHS_LORENZ96, None, float(F))


def rhs_callback(fn) -> Rhs:
    return Rhs(RHS_CALLBACK, None, 0.0, fn)


@dataclass
class Solve:
    t: np.ndarray
    y: np.ndarray  # (n_y_out, n)
    stats: OracleStats
    trace: list


def solve_vector(integrator: str, rhs: Rhs, y0, tspan, options: OracleOptions | None = None, trace: bool = False, trace_cap: int = 1 << 20,
                 max_steps: int = 0) -> Solve:
    y0 = _f64(y0)
    tspan = _f64(tspan)
    n = y0.size
    options = options or new_options()
    t_out = np.full(tspan.size, np.nan)  # the time list can be shorter than tspan (tStart repeated)
    y_out = np.empty((tspan.size, n), dtype=np.float64)
    n_y = C.c_size_t(0)
    st = OracleStats()
    tr = (StepRecord * trace_cap)() if trace else None
    n_tr = C.c_size_t(0)
    cb, keep = rhs._cb()
    rc = lib().oracle_solve_vector(
        integrator.encode(), C.c_int(rhs.kind), _p(rhs.param) if rhs.param is not None else None, C.c_double(rhs.scalar),
        cb, None, C.c_size_t(n), _p(y0), _p(tspan), C.c_size_t(tspan.size), C.byref(options), _p(t_out), _p(y_out),
        C.c_size_t(tspan.size), C.byref(n_y), C.byref(st), tr, C.c_size_t(trace_cap if trace else 0), C.byref(n_tr), C.c_long(max_steps))
    del keep
    _check(rc)
    recs = [(tr[i].t, tr[i].dt_used, tr[i].error, tr[i].attempts) for i in range(min(n_tr.value, trace_cap))] if trace else []
    return Solve(t_out[~np.isnan(t_out)], y_out[: n_y.value].copy(), st, recs)


def solve_scalar(integrator: str, y0: float, tspan, options: OracleOptions | None = None, rhs_scale_c: float = -0.1, callback=None):
    tspan = _f64(tspan)
    options = options or new_options()
    t_out = np.full(tspan.size, np.nan)
    y_out = np.empty(tspan.size)
    n_y = C.c_size_t(0)
    st = OracleStats()
    if callback is not None:
        r = rhs_callback(lambda t, y: np.array([callback(t, float(y[0]))]))
        cb, keep = r._cb()
    else:
        cb, keep = C.cast(None, RHS_CB), None
    rc = lib().oracle_solve_scalar(integrator.encode(), C.c_double(rhs_scale_c), cb, None, C.c_double(y0), _p(tspan),
                                   C.c_size_t(tspan.size), C.byref(options), _p(t_out), _p(y_out), C.byref(n_y), C.byref(st))
    del keep
    _check(rc)
    return t_out[~np.isnan(t_out)], y_out[: n_y.value].copy(), st


def step_vector(integrator: str, rhs: Rhs, t: float, y, fsal, dt: float, options: OracleOptions | None = None):
    """One IntegratorProc call: returns (yNew, newFSAL, dtUsed, error, stats)."""
    y = _f64(y)
    fsal = _f64(fsal)
    n = y.size
    options = options or new_options()
    y_new = np.empty(n)
    f_new = np.empty(n)
    dt_used = C.c_double(0)
    err = C.c_double(0)
    st = OracleStats()
    cb, keep = rhs._cb()
    rc = lib().oracle_step_vector(integrator.encode(), C.c_int(rhs.kind), _p(rhs.param) if rhs.param is not None else None,
                                  C.c_double(rhs.scalar), cb, None, C.c_size_t(n), C.c_double(t), _p(y), _p(fsal), C.c_double(dt),
                                  C.byref(options), _p(y_new), _p(f_new), C.byref(dt_used), C.byref(err), C.byref(st))
    del keep
    _check(rc)
    return y_new, f_new, dt_used.value, err.value, st


def _kptrs(ks):
    ks = [_f64(k) for k in ks]
    arr = (C.c_void_p * len(ks))(*[k.ctypes.data for k in ks])
    return ks, arr


def weighted_stage(w, c: float, y, ks) -> np.ndarray:
    y = _f64(y)
    w = _f64(w)
    ks, arr = _kptrs(ks)
    out = np.empty_like(y)
    _check(lib().oracle_weighted_stage(C.c_size_t(y.size), C.c_int(len(ks)), _p(w), C.c_double(c), _p(y), arr, _p(out)))
    return out


def pair_stage_input(method: str, s: int, dt: float, y, ks) -> np.ndarray:
    y = _f64(y)
    ks, arr = _kptrs(ks)
    out = np.empty_like(y)
    _check(lib().oracle_pair_stage_input(method.encode(), C.c_int(s), C.c_size_t(y.size), C.c_double(dt), _p(y), arr, _p(out)))
    return out


def pair_finish(method: str, dt: float, absTol: float, relTol: float, y, ks):
    """Returns (yNew, error_y, sum(err1^2), error) of ode.nim:301-303 + 61-65."""
    y = _f64(y)
    ks, arr = _kptrs(ks)
    y_new = np.empty_like(y)
    e_y = np.empty_like(y)
    S = C.c_double(0)
    E = C.c_double(0)
    _check(lib().oracle_pair_finish(method.encode(), C.c_size_t(y.size), C.c_double(dt), C.c_double(absTol), C.c_double(relTol),
                                    _p(y), arr, _p(y_new), _p(e_y), C.byref(S), C.byref(E)))
    return y_new, e_y, S.value, E.value


def rk4_stage_input(cfac: float, dt: float, y, k) -> np.ndarray:
    y = _f64(y)
    k = _f64(k)
    out = np.empty_like(y)
    _check(lib().oracle_rk4_stage_input(C.c_size_t(y.size), C.c_double(cfac), C.c_double(dt), _p(y), _p(k), _p(out)))
    return out


def rk4_combine(dt: float, y, k1, k2, k3, k4) -> np.ndarray:
    y, k1, k2, k3, k4 = map(_f64, (y, k1, k2, k3, k4))
    out = np.empty_like(y)
    _check(lib().oracle_rk4_combine(C.c_size_t(y.size), C.c_double(dt), _p(y), _p(k1), _p(k2), _p(k3), _p(k4), _p(out)))
    return out


def hermite(x: float, x1: float, x2: float, y1, y2, dy1, dy2) -> np.ndarray:
    y1, y2, dy1, dy2 = map(_f64, (y1, y2, dy1, dy2))
    out = np.empty_like(y1)
    _check(lib().oracle_hermite(C.c_size_t(y1.size), C.c_double(x), C.c_double(x1), C.c_double(x2), _p(y1), _p(y2), _p(dy1), _p(dy2), _p(out)))
    return out


def rhs_eval(rhs: Rhs, t: float, y) -> np.ndarray:
    y = _f64(y)
    out = np.empty_like(y)
    _check(lib().oracle_rhs_eval(C.c_int(rhs.kind), _p(rhs.param) if rhs.param is not None else None, C.c_double(rhs.scalar),
                                 C.c_size_t(y.size), C.c_double(t), _p(y), _p(out)))
    return out


def vector_binop(op: int, a, b) -> np.ndarray:
    a, b = _f64(a), _f64(b)
    out = np.empty(max(a.size, b.size))
    _check(lib().oracle_vector_binop(C.c_int(op), _p(a), C.c_size_t(a.size), _p(b), C.c_size_t(b.size), _p(out)))
    return out[: a.size]


def vector_unop(op: int, d: float, a) -> np.ndarray:
    a = _f64(a)
    out = np.empty_like(a)
    _check(lib().oracle_vector_unop(C.c_int(op), C.c_double(d), _p(a), C.c_size_t(a.size), _p(out)))
    return out


def vector_sum(a) -> float:
    a = _f64(a)
    return lib().oracle_vector_sum(_p(a), a.size)


def vector_norm(a, p: int = 2) -> float:
    a = _f64(a)
    return lib().oracle_vector_norm(_p(a), a.size, p)


def vector_dot(a, b) -> float:
    a, b = _f64(a), _f64(b)
    return lib().oracle_vector_dot(_p(a), _p(b), a.size)


def is_close(a, b, tol: float = 1e-3) -> bool:
    if np.isscalar(a):
        return bool(lib().oracle_is_close_scalar(float(a), float(b), float(tol)))
    a, b = _f64(a), _f64(b)
    return bool(lib().oracle_is_close_vec(_p(a), _p(b), a.size, float(tol)))


def pair_tableau(method: str) -> dict:
    stages, order, n_b, n_bhat, direct = (C.c_int(0) for _ in range(5))
    c = np.zeros(10)
    a = np.zeros((10, 9))
    b = np.zeros(9)
    bhat = np.zeros(9)
    rc = lib().oracle_pair_tableau(method.encode(), C.byref(stages), C.byref(order), C.byref(n_b), C.byref(n_bhat), C.byref(direct),
                                   _p(c), _p(a), _p(b), _p(bhat))
    if rc:
        raise ValueError_(method)
    return dict(stages=stages.value, order=order.value, n_b=n_b.value, n_bhat=n_bhat.value, err_direct=bool(direct.value),
                c=c, a=a, b=b, bhat=bhat)


# ----------------------------------------------------------------------------------------------------
# Hermite interpolation of a data set + cumulative quadrature (quad_oracle.hpp)
# ----------------------------------------------------------------------------------------------------
FN_CB = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.c_size_t, C.c_void_p)


class Defect(RuntimeError):
    """Stands for Nim's AssertionDefect / IndexDefect (not a ValueError)."""


def _check_q(rc: int):
    if rc == 3:
        raise Defect(lib().oracle_last_error().decode())
    _check(rc)


def _seq(a, scalar: bool):
    """Sequence of T as a contiguous (m, n) array; n = 1 for scalars."""
    a = _f64(a)
    return a.reshape(-1, 1) if scalar else np.atleast_2d(a)


def _out(res, n_out, scalar):
    r = res[: n_out.value].copy()
    return r[:, 0] if scalar else r


def hermite_interpolate(x, t, y, dy, scalar: bool = False) -> np.ndarray:
    """utils.nim:282-312. y, dy: (len(t), n) arrays (or 1-D with scalar=True). Returns (n_out, n)."""
    x, t = _f64(x), _f64(t)
    y, dy = _seq(y, scalar), _seq(dy, scalar)
    n = y.shape[1]
    res = np.empty((max(x.size, 1), n))
    n_out = C.c_size_t(0)
    _check_q(lib().oracle_hermite_interpolate(C.c_int(scalar), C.c_size_t(n), _p(x), C.c_size_t(x.size), _p(t), C.c_size_t(t.size), _p(y), _p(dy),
                                              _p(res), C.byref(n_out)))
    return _out(res, n_out, scalar)


def _cum(fn_name, Y, X, scalar):
    X = _f64(X)
    Y = _seq(Y, scalar)
    n = Y.shape[1]
    res = np.empty((max(X.size, 1), n))
    n_out = C.c_size_t(0)
    _check_q(getattr(lib(), fn_name)(C.c_int(scalar), C.c_size_t(n), _p(Y), _p(X), C.c_size_t(X.size), _p(res), C.byref(n_out)))
    return _out(res, n_out, scalar)


def cumtrapz(Y, X, scalar: bool = False) -> np.ndarray:
    """integrate.nim:119-135."""
    return _cum("oracle_cumtrapz", Y, X, scalar)


def cumsimpson(Y, X, scalar: bool = False) -> np.ndarray:
    """integrate.nim:330-378."""
    return _cum("oracle_cumsimpson", Y, X, scalar)


def _cum_fn(fn_name, f, X, dx, n, scalar):
    X = _f64(X)
    res = np.empty((max(X.size, 1), n))
    n_out = C.c_size_t(0)
    evals = C.c_long(0)

    def tramp(t, op, nn, _user):
        out = np.ctypeslib.as_array(op, shape=(nn,))
        out[:] = f(t)

    cb = FN_CB(tramp)
    _check_q(getattr(lib(), fn_name)(C.c_int(scalar), C.c_size_t(n), cb, None, _p(X), C.c_size_t(X.size), C.c_double(dx), _p(res), C.byref(n_out),
                                     C.byref(evals)))
    return _out(res, n_out, scalar), evals.value


def cumtrapz_fn(f, X, dx: float = 1e-5, n: int = 1, scalar: bool = False):
    """integrate.nim:138-175; f(t) -> array of n values. Returns (values, number of evaluations of f)."""
    return _cum_fn("oracle_cumtrapz_fn", f, X, dx, n, scalar)


def cumsimpson_fn(f, X, dx: float = 1e-5, n: int = 1, scalar: bool = False):
    """integrate.nim:379-400."""
    return _cum_fn("oracle_cumsimpson_fn", f, X, dx, n, scalar)


def simpson_weights(h1: float, h2: float, tail: bool = False):
    a, b, e = C.c_double(0), C.c_double(0), C.c_double(0)
    lib().oracle_simpson_weights(C.c_int(tail), C.c_double(h1), C.c_double(h2), C.byref(a), C.byref(b), C.byref(e))
    return a.value, b.value, e.value


# ----------------------------------------------------------------------------------------------------
# courtesy baseline: fused, multi-threaded CPU code for the diag-linear IVP (NOT the reference's behaviour)
# ----------------------------------------------------------------------------------------------------
class FusedStats(C.Structure):
    _fields_ = [("steps", C.c_long), ("attempts", C.c_long), ("rejected", C.c_long), ("limiter_hits", C.c_long), ("seconds", C.c_double),
                ("threads", C.c_int), ("_pad", C.c_int)]


def fused_mt_solve_diag(integrator: str, lam, y0, t_end: float, options: OracleOptions | None = None, max_steps: int = 0, threads: int = 0):
    """One fused pass per attempt over all host cores (std::thread chunks). Returns (y_end, FusedStats)."""
    lam, y = _f64(lam), _f64(y0).copy()
    st = FusedStats()
    options = options or new_options()
    _check(lib().oracle_fused_mt_solve_diag(integrator.encode(), _p(lam), _p(y), C.c_size_t(y.size), C.c_double(t_end), C.byref(options),
                                            C.c_long(max_steps), C.c_int(threads), C.byref(st)))
    return y, st
