"""numericalnim_b200 — B200-native explicit Runge–Kutta stepper behind numericalnim's solveODE interface.

Scope: ONE hot path of SciNim/numericalnim (ode.nim's RK stage loop + adaptive error norm + the Vector[T]
ops underneath) as hand-written sm_100a CUDA kernels in ``lib/libb200rk.so`` (C-ABI: ``include/b200rk.h``).
This package is the host-side mirror of the reference interface over that C-ABI; it has no CPU fallback.
"""
from ._capi import B200rkError, B200rkValueError, LIB_PATH, SYMBOLS  # noqa: F401
from .ode import (  # noqa: F401
    Context, GpuVector, JitRhs, NumContext, ODEoptions, Solver, adaptiveODE, allODE, combineErr, default_context, fixedODE,
    cumsimpson, cumtrapz, hermiteInterpolate, hermiteSpline, integratorStep, linspace, newNumContext, newODEoptions, newVector, rhsDiagLinear, rhsJit, rhsJitStencil, JitStencilRhs, jitStencilCompileOnly, rhsLorenz96,
    rhsScale, jitCompileOnly, rk4Combine, set_default_context, solveODE, stageAccum,
)

__all__ = [
    "solveODE", "newODEoptions", "ODEoptions", "GpuVector", "newVector", "NumContext", "newNumContext", "fixedODE",
    "adaptiveODE", "allODE", "linspace", "hermiteSpline", "integratorStep", "Solver", "Context", "default_context",
    "set_default_context", "rhsScale", "rhsDiagLinear", "rhsLorenz96", "rhsJit", "JitRhs", "jitCompileOnly", "rhsJitStencil", "JitStencilRhs", "jitStencilCompileOnly", "hermiteInterpolate", "cumtrapz", "cumsimpson", "stageAccum", "combineErr", "rk4Combine",
]
