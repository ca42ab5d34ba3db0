"""ctypes declarations for libb200rk.so — one entry per symbol of include/b200rk.h.

There is no CPU fallback: if the CUDA library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200rk.so")

OK, EINVAL, ECUDA, ENCCL, ENOMEM, ECALLBACK, ENONFINITE = range(7)

METHODS = ("dopri54", "tsit54", "vern65", "rk4", "rk21", "bs32", "heun2", "ralston2", "kutta3", "heun3",
           "ralston3", "ssprk3", "ralston4", "kutta4")  # enum b200rk_method order
RHS_SCALE, RHS_DIAG_LINEAR, RHS_LORENZ96 = 0, 1, 2
K_STAGE, K_FINISH, K_RHS, K_OTHER, K_FUSED, K_QUAD, K_COUNT = 0, 1, 2, 3, 4, 5, 6


class Options(C.Structure):
    """b200rk_options == ODEoptions (ode.nim:26-34)."""
    _fields_ = [(n, C.c_double) for n in ("dt", "dtMax", "dtMin", "tStart", "absTol", "relTol", "scaleMax", "scaleMin")]

    def __repr__(self):
        return "ODEoptions(" + ", ".join(f"{n}={getattr(self, n)!r}" for n, _ in self._fields_) + ")"


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("steps", "attempts", "rejected", "limiter_hits", "rhs_evals", "launches", "collectives")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Profile(C.Structure):
    _fields_ = [("launches", C.c_int64 * K_COUNT), ("ms", C.c_double * K_COUNT), ("algorithmic_bytes", C.c_double * K_COUNT)]


RHS_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p)
FN_OF_T = C.CFUNCTYPE(C.c_int, C.c_double, C.c_void_p, C.c_void_p)

_DECLS = {
    # name: (restype, argtypes)
    "b200rk_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "b200rk_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "b200rk_init_distributed": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b200rk_destroy": (None, [C.c_void_p]),
    "b200rk_last_error": (C.c_char_p, [C.c_void_p]),
    "b200rk_stream": (C.c_void_p, [C.c_void_p]),
    "b200rk_synchronize": (C.c_int, [C.c_void_p]),
    "b200rk_rank": (C.c_int, [C.c_void_p]),
    "b200rk_world": (C.c_int, [C.c_void_p]),
    "b200rk_set": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "b200rk_get": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "b200rk_profile_reset": (C.c_int, [C.c_void_p]),
    "b200rk_profile_read": (C.c_int, [C.c_void_p, C.POINTER(Profile)]),
    "b200rk_ctx_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "b200rk_options_new": (C.c_int, [C.POINTER(Options)] + [C.c_double] * 8),
    "b200rk_options_default": (None, [C.POINTER(Options)]),
    "b200rk_method_from_name": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "b200rk_method_name": (C.c_char_p, [C.c_int]),
    "b200rk_method_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "b200rk_method_tableau": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200rk_shard_range": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "b200rk_vec_new": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "b200rk_vec_free": (C.c_int, [C.c_void_p]),
    "b200rk_vec_len": (C.c_size_t, [C.c_void_p]),
    "b200rk_vec_local_len": (C.c_size_t, [C.c_void_p]),
    "b200rk_vec_local_offset": (C.c_size_t, [C.c_void_p]),
    "b200rk_vec_data": (C.c_void_p, [C.c_void_p]),
    "b200rk_vec_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200rk_vec_download": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200rk_vec_upload_local": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200rk_vec_download_local": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200rk_vec_upload_local_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200rk_vec_copy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200rk_vec_fill": (C.c_int, [C.c_void_p, C.c_double]),
    "b200rk_vec_add": (C.c_int, [C.c_void_p] * 3),
    "b200rk_vec_sub": (C.c_int, [C.c_void_p] * 3),
    "b200rk_vec_hmul": (C.c_int, [C.c_void_p] * 3),
    "b200rk_vec_hdiv": (C.c_int, [C.c_void_p] * 3),
    "b200rk_vec_scale": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "b200rk_vec_div_scalar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double]),
    "b200rk_vec_add_scalar": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "b200rk_vec_neg": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200rk_vec_abs": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200rk_vec_sum": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "b200rk_hermite": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 4),
    "b200rk_builtin_rhs_new": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.POINTER(RHS_FN), C.POINTER(C.c_void_p)]),
    "b200rk_builtin_rhs_free": (C.c_int, [C.c_void_p]),
    "b200rk_jit_rhs_new": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(RHS_FN), C.POINTER(C.c_void_p)]),
    "b200rk_jit_stencil_rhs_new": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(RHS_FN), C.POINTER(C.c_void_p)]),
    "b200rk_jit_stencil_compile_only": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]),
    "b200rk_jit_rhs_set_scalars": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "b200rk_jit_rhs_free": (C.c_int, [C.c_void_p]),
    "b200rk_jit_compile_only": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]),
    "b200rk_hermite_interpolate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                             C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200rk_cumtrapz": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200rk_cumsimpson": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200rk_cumtrapz_fn": (C.c_int, [C.c_void_p, FN_OF_T, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200rk_cumsimpson_fn": (C.c_int, [C.c_void_p, FN_OF_T, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200rk_hermite_plan": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]),
    "b200rk_simpson_weights": (C.c_int, [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200rk_step": (C.c_int, [C.c_void_p, C.c_int, RHS_FN, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double,
                              C.POINTER(Options), C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200rk_solve": (C.c_int, [C.c_void_p, C.c_int, RHS_FN, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Options),
                               C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(Stats)]),
    "b200rk_solve_host": (C.c_int, [C.c_void_p, C.c_int, RHS_FN, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.POINTER(Options), C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(Stats)]),
    "b200rk_solver_new": (C.c_int, [C.c_void_p, C.c_int, RHS_FN, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "b200rk_solver_advance": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "b200rk_solver_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_void_p)]),
    "b200rk_solver_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "b200rk_solver_free": (C.c_int, [C.c_void_p]),
    "b200rk_stage_accum": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "b200rk_combine_err": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_void_p),
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200rk_rk4_combine": (C.c_int, [C.c_void_p, C.c_double] + [C.c_void_p] * 6),
}

SYMBOLS = tuple(_DECLS)
_lib = None


class B200rkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200rk error {code}: {msg}")
        self.code = code


class B200rkValueError(ValueError):
    """B200RK_EINVAL — the role Nim's ValueError plays in the reference."""


def lib():
    """Load libb200rk.so (built in-tree by __graft_entry__.build() / `make -C numericalnim_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). numericalnim_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _DECLS.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, ctx=None):
    if rc == OK:
        return
    msg = lib().b200rk_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc == EINVAL:
        raise B200rkValueError(msg)
    raise B200rkError(rc, msg)
