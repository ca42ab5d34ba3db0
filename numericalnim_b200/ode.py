"""Host-side mirror of numericalnim's ODE interface over the C-ABI (include/b200rk.h).

Names, argument meaning and error behaviour follow the reference so tests read like its own:

    solveODE(f, y0, tspan, options=DEFAULT_ODEoptions, ctx=None, integrator="dopri54")   ode.nim:589-591
    newODEoptions(dt, absTol, relTol, dtMax, dtMin, scaleMax, scaleMin, tStart)          ode.nim:78-79
    ODEProc:  f(t, y, ctx) -> dy                                                          ode.nim:36
    fixedODE / adaptiveODE / allODE                                                       ode.nim:40-42
    GpuVector  (device-resident stand-in for Vector[float], utils.nim:14-271)
    NumContext (commonTypes.nim:3-39)

All arithmetic runs in libb200rk.so on the GPU; nothing here computes on the CPU and there is no
fallback. ValueError is raised where the reference raises ValueError.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Sequence

import numpy as np

from . import _capi as capi
from ._capi import Options as ODEoptions  # noqa: N814  (reference spelling)

fixedODE = ["heun2", "ralston2", "kutta3", "heun3", "ralston3", "ssprk3", "ralston4", "kutta4", "rk4"]  # ode.nim:40
adaptiveODE = ["rk21", "bs32", "dopri54", "tsit54", "vern65"]  # ode.nim:41
allODE = fixedODE + adaptiveODE  # ode.nim:42


# ----------------------------------------------------------------------------------------------------
# context
# ----------------------------------------------------------------------------------------------------
class Context:
    """b200rk_ctx: one GPU, one stream, optional NCCL communicator (one process per GPU)."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_id: bytes | None = None):
        L = capi.lib()
        h = C.c_void_p()
        if world > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("distributed context needs the 128-byte NCCL unique id")
            buf = C.create_string_buffer(nccl_id, 128)
            capi.check(L.b200rk_init_distributed(C.byref(h), device, rank, world, buf))
        else:
            capi.check(L.b200rk_init(C.byref(h), device))
        self._h = h
        self.device, self.rank, self.world = device, rank, world

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().b200rk_nccl_unique_id(buf))
        return buf.raw

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        return capi.lib().b200rk_stream(self._h) or 0

    def synchronize(self):
        capi.check(capi.lib().b200rk_synchronize(self._h), self._h)

    def set(self, key: str, value: int):
        capi.check(capi.lib().b200rk_set(self._h, key.encode(), int(value)), self._h)

    def get(self, key: str) -> int:
        v = C.c_int64(0)
        capi.check(capi.lib().b200rk_get(self._h, key.encode(), C.byref(v)), self._h)
        return v.value

    def profile_reset(self):
        capi.check(capi.lib().b200rk_profile_reset(self._h), self._h)

    def profile_read(self) -> dict:
        p = capi.Profile()
        capi.check(capi.lib().b200rk_profile_read(self._h, C.byref(p)), self._h)
        names = ("stage", "finish", "rhs", "other", "fused", "quad")
        return {n: dict(launches=int(p.launches[i]), ms=float(p.ms[i]), bytes=float(p.algorithmic_bytes[i])) for i, n in enumerate(names)}

    def stats(self) -> dict:
        s = capi.Stats()
        capi.check(capi.lib().b200rk_ctx_stats(self._h, C.byref(s)), self._h)
        return s.as_dict()

    def close(self):
        if self._h:
            capi.lib().b200rk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Context | None = None


def default_context() -> Context:
    """Process-wide context on cuda:LOCAL_RANK (single GPU). Distributed runs build their own Context."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def set_default_context(ctx: Context | None):
    global _default_ctx
    _default_ctx = ctx


# ----------------------------------------------------------------------------------------------------
# GpuVector
# ----------------------------------------------------------------------------------------------------
class GpuVector:
    """Device-resident Vector[float]. Operators allocate a fresh result, like the reference's
    (utils.nim:59-64); each is one kernel launch. The solver never goes through these — it uses the
    fused stage / finish kernels — they exist so user right-hand sides can be written as in the
    reference (``-0.1 * y``)."""

    __slots__ = ("ctx", "_h", "_owned")

    def __init__(self, ctx: Context, handle, owned: bool = True):
        self.ctx, self._h, self._owned = ctx, handle, owned

    # -- construction / transfer
    @classmethod
    def empty(cls, n: int, ctx: Context | None = None) -> "GpuVector":
        ctx = ctx or default_context()
        h = C.c_void_p()
        capi.check(capi.lib().b200rk_vec_new(ctx.handle, n, C.byref(h)), ctx.handle)
        return cls(ctx, h)

    @classmethod
    def from_host(cls, components, ctx: Context | None = None) -> "GpuVector":
        """newVector(components) (utils.nim:19-20): copies the GLOBAL host array (each rank its slice)."""
        a = np.ascontiguousarray(np.asarray(components, dtype=np.float64))
        v = cls.empty(a.size, ctx)
        capi.check(capi.lib().b200rk_vec_upload(v._h, a.ctypes.data), v.ctx.handle)
        return v

    @classmethod
    def from_local(cls, n_global: int, local, ctx: Context | None = None) -> "GpuVector":
        a = np.ascontiguousarray(np.asarray(local, dtype=np.float64))
        v = cls.empty(n_global, ctx)
        if a.size != v.local_len:
            raise ValueError(f"local shard has {a.size} elements, expected {v.local_len}")
        capi.check(capi.lib().b200rk_vec_upload_local(v._h, a.ctypes.data), v.ctx.handle)
        return v

    def to_numpy(self) -> np.ndarray:
        """Global-length host copy; in a sharded context only this rank's slice is filled (rest NaN)."""
        out = np.full(len(self), np.nan) if self.ctx.world > 1 else np.empty(len(self))
        capi.check(capi.lib().b200rk_vec_download(self._h, out.ctypes.data), self.ctx.handle)
        return out

    def local_numpy(self) -> np.ndarray:
        out = np.empty(self.local_len)
        capi.check(capi.lib().b200rk_vec_download_local(self._h, out.ctypes.data), self.ctx.handle)
        return out

    @property
    def components(self) -> np.ndarray:  # Vector.components (utils.nim:16)
        return self.to_numpy()

    def __len__(self):
        return capi.lib().b200rk_vec_len(self._h)

    @property
    def len(self):  # Vector.len (utils.nim:17)
        return len(self)

    @property
    def local_len(self) -> int:
        return capi.lib().b200rk_vec_local_len(self._h)

    @property
    def local_offset(self) -> int:
        return capi.lib().b200rk_vec_local_offset(self._h)

    @property
    def data_ptr(self) -> int:
        return capi.lib().b200rk_vec_data(self._h) or 0

    @property
    def __cuda_array_interface__(self):  # zero-copy view for torch.as_tensor(..., device="cuda") / cupy
        return dict(shape=(self.local_len,), typestr="<f8", data=(self.data_ptr, False), version=3, stream=self.ctx.stream or None)

    def free(self):
        # a vector that outlives its (explicitly closed) context must not reach into it: b200rk_vec_free returns the
        # buffer to the context's pool
        if self._owned and self._h and self.ctx._h:
            capi.lib().b200rk_vec_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- operators (utils.nim:59-223)
    def _new_like(self) -> "GpuVector":
        return GpuVector.empty(len(self), self.ctx)

    def _bin(self, fn, other) -> "GpuVector":
        out = self._new_like()
        capi.check(fn(out._h, self._h, other._h), self.ctx.handle)
        return out

    def __add__(self, o):
        L = capi.lib()
        if isinstance(o, GpuVector):
            return self._bin(L.b200rk_vec_add, o)
        out = self._new_like()
        capi.check(L.b200rk_vec_add_scalar(out._h, float(o), self._h), self.ctx.handle)
        return out

    __radd__ = __add__

    def __sub__(self, o):
        L = capi.lib()
        if isinstance(o, GpuVector):
            return self._bin(L.b200rk_vec_sub, o)
        out = self._new_like()
        capi.check(L.b200rk_vec_add_scalar(out._h, -float(o), self._h), self.ctx.handle)
        return out

    def __mul__(self, d):
        if isinstance(d, GpuVector):  # Vector * Vector is the dot product (utils.nim:181-185)
            return self.hmul(d).sum()
        out = self._new_like()
        capi.check(capi.lib().b200rk_vec_scale(out._h, float(d), self._h), self.ctx.handle)
        return out

    __rmul__ = __mul__

    def __truediv__(self, d):
        out = self._new_like()
        capi.check(capi.lib().b200rk_vec_div_scalar(out._h, self._h, float(d)), self.ctx.handle)
        return out

    def __neg__(self):
        out = self._new_like()
        capi.check(capi.lib().b200rk_vec_neg(out._h, self._h), self.ctx.handle)
        return out

    def __abs__(self):
        out = self._new_like()
        capi.check(capi.lib().b200rk_vec_abs(out._h, self._h), self.ctx.handle)
        return out

    def hmul(self, o: "GpuVector") -> "GpuVector":  # `*.`
        return self._bin(capi.lib().b200rk_vec_hmul, o)

    def hdiv(self, o: "GpuVector") -> "GpuVector":  # `/.`
        return self._bin(capi.lib().b200rk_vec_hdiv, o)

    def sum(self) -> float:
        out = C.c_double(0)
        capi.check(capi.lib().b200rk_vec_sum(self._h, C.byref(out)), self.ctx.handle)
        return out.value

    def clone(self) -> "GpuVector":
        out = self._new_like()
        capi.check(capi.lib().b200rk_vec_copy(out._h, self._h), self.ctx.handle)
        return out

    def copy_from(self, src: "GpuVector"):
        capi.check(capi.lib().b200rk_vec_copy(self._h, src._h), self.ctx.handle)


def newVector(components, ctx: Context | None = None) -> GpuVector:  # utils.nim:19
    return GpuVector.from_host(components, ctx)


def hermiteSpline(x: float, x1: float, x2: float, y1: GpuVector, y2: GpuVector, dy1: GpuVector, dy2: GpuVector) -> GpuVector:
    """utils.nim:273-279 on device vectors."""
    out = y1._new_like()
    capi.check(capi.lib().b200rk_hermite(out._h, x, x1, x2, y1._h, y2._h, dy1._h, dy2._h), y1.ctx.handle)
    return out


def _adopt_list(ctx: "Context", slots, n_out) -> list:
    return [GpuVector(ctx, C.c_void_p(slots[i])) for i in range(n_out.value)]


def hermiteInterpolate(x: Sequence[float], t: Sequence[float], y: Sequence["GpuVector"], dy: Sequence["GpuVector"]) -> list:
    """utils.nim:282-312 on a trajectory of device vectors: the cubic Hermite spline through (t[i], y[i]) with slopes
    dy[i], evaluated at every x in ONE kernel. Like the reference: sorted x silently drops samples outside the data,
    unsorted x raises ValueError for them."""
    y, dy = list(y), list(dy)
    if len(y) != len(t) or len(dy) != len(t):
        raise ValueError("t, y and dy must have the same length")
    if not y:
        raise ValueError("index out of bounds, the container is empty")
    ctx = y[0].ctx
    xa = np.ascontiguousarray(np.asarray(list(x), dtype=np.float64))
    ta = np.ascontiguousarray(np.asarray(list(t), dtype=np.float64))
    slots = (C.c_void_p * max(xa.size, 1))()
    n_out = C.c_size_t(0)
    capi.check(capi.lib().b200rk_hermite_interpolate(ctx.handle, xa.ctypes.data, xa.size, ta.ctypes.data, ta.size, _ptr_array(y), _ptr_array(dy),
                                                     slots, C.byref(n_out)), ctx.handle)
    return _adopt_list(ctx, slots, n_out)


class _PyFnOfT:
    """Adapts a Python integrand ``f(x, ctx) -> GpuVector`` (NumContextProc[T, float]) to b200rk_fn_of_t."""

    def __init__(self, f: Callable, dctx: "Context", numctx):
        self.exc = None
        self.calls = 0

        def tramp(t, out_h, _user):
            try:
                self.calls += 1
                out = GpuVector(dctx, C.c_void_p(out_h), owned=False)
                r = f(t, numctx)
                if not isinstance(r, GpuVector):
                    raise TypeError("the integrand must return a GpuVector")
                if r._h.value != out_h:
                    out.copy_from(r)
                return 0
            except BaseException as e:  # noqa: BLE001 — must not unwind through C
                self.exc = e
                return 1

        self.fn = capi.FN_OF_T(tramp)


def _cumulative(name: str, first, X, ctx, dx, like):
    L = capi.lib()
    Xa = np.ascontiguousarray(np.asarray(list(X), dtype=np.float64))
    n_out = C.c_size_t(0)
    slots = (C.c_void_p * max(Xa.size, 1))()
    if callable(first):  # cumtrapz(f, X, ctx, dx) / cumsimpson(f, X, ctx, dx)
        numctx = ctx if ctx is not None else newNumContext()
        if like is None:
            like = first(float(Xa.min()) if Xa.size else 0.0, numctx)  # learn T's size (one extra evaluation; pass `like` to avoid it)
        dctx = like.ctx
        cb = _PyFnOfT(first, dctx, numctx)
        rc = getattr(L, name + "_fn")(dctx.handle, cb.fn, None, len(like), Xa.ctypes.data, Xa.size, float(dx), slots, C.byref(n_out))
        if cb.exc is not None:
            raise cb.exc
        capi.check(rc, dctx.handle)
        return _adopt_list(dctx, slots, n_out)
    Y = list(first)  # cumtrapz(Y, X) / cumsimpson(Y, X)
    if len(Y) != Xa.size:
        raise ValueError("X and Y must have the same length")
    if not Y:
        raise ValueError("x is empty!")
    dctx = Y[0].ctx
    capi.check(getattr(L, name)(dctx.handle, _ptr_array(Y), Xa.ctypes.data, Xa.size, slots, C.byref(n_out)), dctx.handle)
    return _adopt_list(dctx, slots, n_out)


def cumtrapz(Y_or_f, X: Sequence[float], ctx: "NumContext | None" = None, dx: float = 1e-5, like: "GpuVector | None" = None) -> list:
    """integrate.nim:119-135 (``cumtrapz(Y, X)``: a list of GpuVector sampled at X) and integrate.nim:138-175
    (``cumtrapz(f, X, ctx, dx)``: ``f(x, ctx) -> GpuVector`` integrated in steps of dx, returned at X)."""
    return _cumulative("b200rk_cumtrapz", Y_or_f, X, ctx, dx, like)


def cumsimpson(Y_or_f, X: Sequence[float], ctx: "NumContext | None" = None, dx: float = 1e-5, like: "GpuVector | None" = None) -> list:
    """integrate.nim:330-378 (``cumsimpson(Y, X)``) and integrate.nim:379-400 (``cumsimpson(f, X, ctx, dx)``)."""
    return _cumulative("b200rk_cumsimpson", Y_or_f, X, ctx, dx, like)


class NumContext:
    """commonTypes.nim:3-39 — mutable bag shared with the user's right-hand side."""

    def __init__(self, fValues: dict | None = None, tValues: dict | None = None):
        self.fValues = dict(fValues or {})
        self.tValues = dict(tValues or {})

    def __getitem__(self, key):
        return self.tValues[str(key)]

    def __setitem__(self, key, val):
        self.tValues[str(key)] = val

    def getF(self, key):
        return self.fValues[str(key)]

    def setF(self, key, val):
        self.fValues[str(key)] = val


def newNumContext(fValues=None, tValues=None) -> NumContext:
    return NumContext(fValues, tValues)


# ----------------------------------------------------------------------------------------------------
# options / helpers
# ----------------------------------------------------------------------------------------------------
def newODEoptions(dt: float = 1e-4, absTol: float = 1e-4, relTol: float = 1e-4, dtMax: float = 1e-2, dtMin: float = 1e-4,
                  scaleMax: float = 4.0, scaleMin: float = 0.1, tStart: float = 0.0) -> ODEoptions:
    """ode.nim:78-102 — raises ValueError on |dtMax| < |dtMin|, |scaleMax| < 1, |scaleMin| > 1."""
    o = ODEoptions()
    capi.check(capi.lib().b200rk_options_new(C.byref(o), dt, absTol, relTol, dtMax, dtMin, scaleMax, scaleMin, tStart))
    return o


def linspace(x1: float, x2: float, N: int) -> list:
    """utils.nim:498-507 (host helper used by the ODE fixtures)."""
    if N <= 0:
        raise ValueError(f"Number of samples {N} must be greater then 0")
    dx = (x2 - x1) / float(N - 1) if N > 1 else float("nan")
    r = [x1]
    for i in range(1, N - 1):
        r.append(x1 + dx * float(i))
    r.append(x2)
    return r


def method_id(integrator: str) -> int:
    m = C.c_int(0)
    capi.check(capi.lib().b200rk_method_from_name(integrator.encode(), C.byref(m)))
    return m.value


# ----------------------------------------------------------------------------------------------------
# right-hand sides
# ----------------------------------------------------------------------------------------------------
class BuiltinRhs:
    """Device-side right-hand side shipped with the library (b200rk_builtin_rhs_new)."""

    def __init__(self, kind: int, scalar: float = 0.0, lam: GpuVector | None = None, ctx: Context | None = None):
        self.ctx = ctx or (lam.ctx if lam is not None else default_context())
        self._lam = lam  # keep alive
        self.fn = capi.RHS_FN()
        self.user = C.c_void_p()
        capi.check(capi.lib().b200rk_builtin_rhs_new(self.ctx.handle, kind, float(scalar), lam._h if lam is not None else None,
                                                     C.byref(self.fn), C.byref(self.user)), self.ctx.handle)

    def __del__(self):
        try:
            if self.user:
                capi.lib().b200rk_builtin_rhs_free(self.user)
        except Exception:
            pass


def rhsScale(c: float, ctx: Context | None = None) -> BuiltinRhs:
    """dy = c*y  (tests/test_ode.nim:5-7 with c = -0.1)."""
    return BuiltinRhs(capi.RHS_SCALE, c, None, ctx)


def rhsDiagLinear(lam: GpuVector) -> BuiltinRhs:
    """dy = -(lambda .* y)."""
    return BuiltinRhs(capi.RHS_DIAG_LINEAR, 0.0, lam)


def rhsLorenz96(F: float = 8.0, ctx: Context | None = None) -> BuiltinRhs:
    """dy[i] = (y[i+1] - y[i-2])*y[i-1] - y[i] + F, cyclic."""
    return BuiltinRhs(capi.RHS_LORENZ96, F, None, ctx)


class JitRhs:
    """Element-local right-hand side given as source (b200rk_jit_rhs_new): ``dy[i] = expr(t, y[i], p0[i].., c0..)``.

    ``expr`` is a CUDA C++ expression in ``t``, ``y``, ``p0..p3`` (the GpuVectors of ``vecs``, one value per
    element) and ``c0..c7`` (``scalars``). NVRTC compiles it INTO the library's fused kernels, so a user-defined
    IVP runs the whole-attempt kernel / the device-resident driver loop / the one-kernel RK4 step exactly like the
    built-in right-hand sides, and through a plain dy = f(t, y) kernel for every other method. Compiled without
    FMA contraction: ``a*b + c`` rounds twice, like the reference's CPU arithmetic. A wrong expression raises
    ValueError carrying the compiler log."""

    def __init__(self, expr: str, vecs: Sequence["GpuVector"] = (), scalars: Sequence[float] = (), ctx: Context | None = None):
        vecs = list(vecs)
        self.ctx = ctx or (vecs[0].ctx if vecs else default_context())
        self._vecs = vecs  # keep alive
        self.expr = expr
        self.fn = capi.RHS_FN()
        self.user = C.c_void_p()
        cs = np.ascontiguousarray(np.asarray(list(scalars), dtype=np.float64))
        capi.check(capi.lib().b200rk_jit_rhs_new(self.ctx.handle, expr.encode(), len(vecs), _ptr_array(vecs) if vecs else None, cs.size,
                                                 cs.ctypes.data if cs.size else None, C.byref(self.fn), C.byref(self.user)), self.ctx.handle)

    def set_scalars(self, scalars: Sequence[float]):
        """New values for c0.. (kernel arguments: no recompilation)."""
        cs = np.ascontiguousarray(np.asarray(list(scalars), dtype=np.float64))
        capi.check(capi.lib().b200rk_jit_rhs_set_scalars(self.user, cs.size, cs.ctypes.data if cs.size else None), self.ctx.handle)

    def __del__(self):
        try:
            if self.user and self.ctx._h:   # the free synchronises the context's stream
                capi.lib().b200rk_jit_rhs_free(self.user)
        except Exception:
            pass


class JitStencilRhs(JitRhs):
    """Stencil right-hand side given as source (b200rk_jit_stencil_rhs_new): ``dy[i] = expr(t, Y(-rl)..Y(+rr), p0[i].., c0..)`` with
    ``Y(d) = y[(i + d) mod N]`` (cyclic). Compiled into a plain dy = f(t, y) kernel for every method and into the one-kernel
    attempt over overlapped tiles for DOPRI54 / Tsit54 / Vern65 (what the built-in Lorenz-96 gets)."""

    def __init__(self, expr: str, radius_left: int, radius_right: int, vecs: Sequence["GpuVector"] = (), scalars: Sequence[float] = (),
                 ctx: Context | None = None):
        vecs = list(vecs)
        self.ctx = ctx or (vecs[0].ctx if vecs else default_context())
        self._vecs = vecs  # keep alive
        self.expr, self.radius_left, self.radius_right = expr, int(radius_left), int(radius_right)
        self.fn = capi.RHS_FN()
        self.user = C.c_void_p()
        cs = np.ascontiguousarray(np.asarray(list(scalars), dtype=np.float64))
        capi.check(capi.lib().b200rk_jit_stencil_rhs_new(self.ctx.handle, expr.encode(), self.radius_left, self.radius_right, len(vecs),
                                                         _ptr_array(vecs) if vecs else None, cs.size, cs.ctypes.data if cs.size else None,
                                                         C.byref(self.fn), C.byref(self.user)), self.ctx.handle)


def rhsJitStencil(expr: str, radius_left: int, radius_right: int, vecs: Sequence["GpuVector"] = (), scalars: Sequence[float] = (),
                  ctx: Context | None = None) -> JitStencilRhs:
    """``rhsJitStencil("((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, scalars=[8.0])`` — see JitStencilRhs."""
    return JitStencilRhs(expr, radius_left, radius_right, vecs, scalars, ctx)


def jitStencilCompileOnly(expr: str, radius_left: int, radius_right: int, n_vec: int = 0, n_scalar: int = 0, pattern: int = -1):
    """Host-only NVRTC compile of one stencil translation unit (no GPU needed): returns (cubin size, log with kernel names)."""
    nb = C.c_size_t(0)
    log = C.create_string_buffer(1 << 16)
    capi.check(capi.lib().b200rk_jit_stencil_compile_only(expr.encode(), radius_left, radius_right, n_vec, n_scalar, pattern, C.byref(nb), log, len(log)))
    return nb.value, log.value.decode()


def rhsJit(expr: str, vecs: Sequence["GpuVector"] = (), scalars: Sequence[float] = (), ctx: Context | None = None) -> JitRhs:
    """``rhsJit("c0*y*(1.0 - y/p0)", vecs=[K], scalars=[r])`` — see JitRhs."""
    return JitRhs(expr, vecs, scalars, ctx)


def jitCompileOnly(expr: str, n_vec: int = 0, n_scalar: int = 0, pattern: int = -1):
    """Host-only NVRTC compile of one translation unit (no GPU needed): returns (cubin bytes, log with kernel names)."""
    L = capi.lib()
    nb = C.c_size_t(0)
    log = C.create_string_buffer(1 << 16)
    capi.check(L.b200rk_jit_compile_only(expr.encode(), n_vec, n_scalar, pattern, None, 0, C.byref(nb), log, len(log)))
    buf = C.create_string_buffer(nb.value)
    capi.check(L.b200rk_jit_compile_only(expr.encode(), n_vec, n_scalar, pattern, buf, nb.value, C.byref(nb), log, len(log)))
    return buf.raw[: nb.value], log.value.decode()


class _PyRhs:
    """Adapts a Python ODEProc ``f(t, y: GpuVector, ctx) -> GpuVector`` to b200rk_rhs_fn. The callable's
    GpuVector operators enqueue kernels on the context stream; the result is copied into ``dydt``."""

    def __init__(self, f: Callable, ctx: Context, numctx: NumContext):
        self.exc = None

        def tramp(t, y_h, dy_h, _user):
            try:
                y = GpuVector(ctx, C.c_void_p(y_h), owned=False)
                dy = GpuVector(ctx, C.c_void_p(dy_h), owned=False)
                r = f(t, y, numctx)
                if not isinstance(r, GpuVector):
                    raise TypeError("right-hand side must return a GpuVector")
                if r._h.value != dy_h:
                    dy.copy_from(r)
                return 0
            except BaseException as e:  # noqa: BLE001 — must not unwind through C
                self.exc = e
                return 1

        self.fn = capi.RHS_FN(tramp)
        self.user = C.c_void_p()


def _resolve_rhs(f, ctx: Context, numctx: NumContext):
    if isinstance(f, (BuiltinRhs, JitRhs)):
        return f
    if callable(f):
        return _PyRhs(f, ctx, numctx)
    raise TypeError("f must be a BuiltinRhs, a JitRhs or a callable f(t, y, ctx)")


# ----------------------------------------------------------------------------------------------------
# solveODE
# ----------------------------------------------------------------------------------------------------
last_stats: dict = {}


def _times(t_out) -> list:
    """The library marks the unused tail of t_out with NaN when the reference's time list is shorter than tspan
    (tStart repeated in tspan, ode.nim:485-487)."""
    return t_out[~np.isnan(t_out)].tolist()


def solveODE(f, y0, tspan: Sequence[float], options: ODEoptions | None = None, ctx: NumContext | None = None,
             integrator: str = "dopri54", device_ctx: Context | None = None):
    """ode.nim:589-651. Returns ``(t, y)``: t is the sorted tspan; y holds one state per returned time.

    y0 may be a GpuVector (device in, list of GpuVector out), a host array / list (host in, list of numpy
    arrays out: the end-to-end path with host<->device copies inside), or a Python float (treated as a
    length-1 vector; floats out) — the three `T` the reference's tests use, minus arraymancer.
    Unknown integrator -> ValueError (ode.nim:650-651)."""
    global last_stats
    L = capi.lib()
    method = method_id(integrator)  # raises ValueError before any work, like the reference's `case`
    options = options if options is not None else newODEoptions()
    numctx = ctx if ctx is not None else newNumContext()  # ode.nim:604-606
    ts = np.ascontiguousarray(np.asarray(list(tspan), dtype=np.float64))
    t_out = np.empty(ts.size)
    st = capi.Stats()
    n_out = C.c_size_t(0)

    if isinstance(y0, GpuVector):
        dctx = y0.ctx
        rhs = _resolve_rhs(f, dctx, numctx)
        slots = (C.c_void_p * max(ts.size, 1))()
        rc = L.b200rk_solve(dctx.handle, method, rhs.fn, rhs.user, y0._h, ts.ctypes.data, ts.size, C.byref(options),
                            t_out.ctypes.data, slots, C.byref(n_out), C.byref(st))
        _reraise(rhs)
        capi.check(rc, dctx.handle)
        last_stats = st.as_dict()
        return _times(t_out), [GpuVector(dctx, C.c_void_p(slots[i])) for i in range(n_out.value)]

    scalar = np.isscalar(y0)
    dctx = device_ctx or default_context()
    y0a = np.ascontiguousarray(np.atleast_1d(np.asarray(y0, dtype=np.float64)))
    rhs = _resolve_rhs(f, dctx, numctx)
    if dctx.world > 1:
        raise ValueError("host-array solveODE is single-GPU; use GpuVector.from_local in sharded runs")
    y_out = np.empty((max(ts.size, 1), y0a.size))
    rc = L.b200rk_solve_host(dctx.handle, method, rhs.fn, rhs.user, y0a.size, y0a.ctypes.data, ts.ctypes.data, ts.size,
                             C.byref(options), t_out.ctypes.data, y_out.ctypes.data, C.byref(n_out), C.byref(st))
    _reraise(rhs)
    capi.check(rc, dctx.handle)
    last_stats = st.as_dict()
    ys = y_out[: n_out.value]
    if scalar:
        return _times(t_out), [float(v[0]) for v in ys]
    return _times(t_out), [v.copy() for v in ys]


def _reraise(rhs):
    exc = getattr(rhs, "exc", None)
    if exc is not None:
        rhs.exc = None
        raise exc


def integratorStep(integrator: str, f, t: float, y: GpuVector, FSAL: GpuVector | None, dt: float,
                   options: ODEoptions | None = None, ctx: NumContext | None = None):
    """One IntegratorProc call (ode.nim:38): returns (yNew, newFSAL, dtUsed, error)."""
    L = capi.lib()
    method = method_id(integrator)
    options = options if options is not None else newODEoptions()
    numctx = ctx if ctx is not None else newNumContext()
    rhs = _resolve_rhs(f, y.ctx, numctx)
    y_new, f_new = y._new_like(), y._new_like()
    dt_used, err = C.c_double(0), C.c_double(0)
    rc = L.b200rk_step(y.ctx.handle, method, rhs.fn, rhs.user, t, y._h, FSAL._h if FSAL is not None else None, dt,
                       C.byref(options), y_new._h, f_new._h, C.byref(dt_used), C.byref(err))
    _reraise(rhs)
    capi.check(rc, y.ctx.handle)
    return y_new, f_new, dt_used.value, err.value


class Solver:
    """Resumable forward driver (b200rk_solver_*): the `while t < tEnd` loop of ODESolver, K steps at a time."""

    def __init__(self, integrator: str, f, y0: GpuVector, t_end: float, options: ODEoptions | None = None, ctx: NumContext | None = None):
        self.dctx = y0.ctx
        self.options = options if options is not None else newODEoptions()
        self.rhs = _resolve_rhs(f, self.dctx, ctx if ctx is not None else newNumContext())
        self._h = C.c_void_p()
        capi.check(capi.lib().b200rk_solver_new(self.dctx.handle, method_id(integrator), self.rhs.fn, self.rhs.user, y0._h,
                                                float(t_end), C.byref(self.options), C.byref(self._h)), self.dctx.handle)

    def advance(self, max_steps: int = -1):
        done, fin = C.c_int64(0), C.c_int(0)
        rc = capi.lib().b200rk_solver_advance(self._h, max_steps, C.byref(done), C.byref(fin))
        _reraise(self.rhs)
        capi.check(rc, self.dctx.handle)
        return done.value, bool(fin.value)

    def state(self):
        t, dt, err, y = C.c_double(0), C.c_double(0), C.c_double(0), C.c_void_p()
        capi.check(capi.lib().b200rk_solver_state(self._h, C.byref(t), C.byref(dt), C.byref(err), C.byref(y)), self.dctx.handle)
        return t.value, dt.value, err.value, GpuVector(self.dctx, y, owned=False)

    def stats(self) -> dict:
        s = capi.Stats()
        capi.check(capi.lib().b200rk_solver_stats(self._h, C.byref(s)), self.dctx.handle)
        return s.as_dict()

    def close(self):
        if self._h:
            if self.dctx._h:   # a solver that outlives its context has nothing left to release
                capi.lib().b200rk_solver_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------
# raw kernels (bandwidth sweep / kernel-level parity)
# ----------------------------------------------------------------------------------------------------
def _ptr_array(vs):
    return (C.c_void_p * len(vs))(*[v._h.value for v in vs])


def stageAccum(w, c: float, y: GpuVector, ks: Sequence[GpuVector], out: GpuVector | None = None, chain: bool = False) -> GpuVector:
    out = out if out is not None else y._new_like()
    wa = np.ascontiguousarray(np.asarray(w, dtype=np.float64))
    capi.check(capi.lib().b200rk_stage_accum(y.ctx.handle, len(ks), wa.ctypes.data, float(c), int(chain), y._h, _ptr_array(ks), out._h), y.ctx.handle)
    return out


def combineErr(integrator: str, dt: float, absTol: float, relTol: float, y: GpuVector, ks: Sequence[GpuVector], want_err_y: bool = False):
    """Returns (yNew, error_y or None, sumsq, error)."""
    y_new = y._new_like()
    e_y = y._new_like() if want_err_y else None
    S, E = C.c_double(0), C.c_double(0)
    capi.check(capi.lib().b200rk_combine_err(y.ctx.handle, method_id(integrator), dt, absTol, relTol, y._h, _ptr_array(ks), y_new._h,
                                             e_y._h if e_y is not None else None, C.byref(S), C.byref(E)), y.ctx.handle)
    return y_new, e_y, S.value, E.value


def rk4Combine(dt: float, y: GpuVector, k1: GpuVector, k2: GpuVector, k3: GpuVector, k4: GpuVector) -> GpuVector:
    out = y._new_like()
    capi.check(capi.lib().b200rk_rk4_combine(y.ctx.handle, dt, y._h, k1._h, k2._h, k3._h, k4._h, out._h), y.ctx.handle)
    return out
