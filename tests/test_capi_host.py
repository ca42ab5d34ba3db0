"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/b200rk.h declares, and its
host-only entry points (options, dispatch, tableaux) behave like the reference. No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import numericalnim_b200 as nn
from numericalnim_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200rk.h")).read()
    declared = set(re.findall(r"B200RK_API\s+[\w\s\*]+?\b(b200rk_\w+)\s*\(", hdr))
    assert len(declared) >= 50
    L = C.CDLL(_capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)


def test_no_cpu_fallback_and_oracle_not_imported():
    """The product package must not import the oracle, and must raise without a CUDA device."""
    import sys
    pkg = os.path.join(ROOT, "numericalnim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src and "rk_oracle" not in src, f
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(_capi.B200rkError):
            nn.Context(0)


def test_options_validation_matches_reference():  # ode.nim:95-102
    with pytest.raises(ValueError, match="dtMin must be less than dtMax"):
        nn.newODEoptions(dtMax=1e-5, dtMin=1e-4)
    with pytest.raises(ValueError, match="scaleMax must be bigger than 1"):
        nn.newODEoptions(scaleMax=0.9)
    with pytest.raises(ValueError, match="scaleMin must be smaller than 1"):
        nn.newODEoptions(scaleMin=-1.1)
    o = nn.newODEoptions(dt=-1e-3, absTol=-1e-5, relTol=-1e-6, dtMax=-1.0, dtMin=-1e-6, scaleMax=-4, scaleMin=-0.1, tStart=-2.0)
    assert (o.dt, o.absTol, o.relTol, o.dtMax, o.dtMin, o.scaleMax, o.scaleMin, o.tStart) == (1e-3, 1e-5, 1e-6, 1.0, 1e-6, 4.0, 0.1, -2.0)
    d = nn.newODEoptions()  # DEFAULT_ODEoptions, ode.nim:78-79,104
    assert (d.dt, d.absTol, d.relTol, d.dtMax, d.dtMin, d.scaleMax, d.scaleMin, d.tStart) == (1e-4, 1e-4, 1e-4, 1e-2, 1e-4, 4.0, 0.1, 0.0)


def test_dispatch_table_matches_reference():  # ode.nim:40-42, 607-651, SURVEY Appendix D
    assert sorted(nn.allODE) == sorted(_capi.METHODS)
    expect = {  # name: (stages, useFSAL, order, adaptive)
        "dopri54": (7, 1, 5.0, 1), "tsit54": (7, 1, 5.0, 1), "vern65": (9, 1, 6.0, 1), "rk4": (4, 0, 4.0, 0),
        "rk21": (2, 0, 2.0, 1), "bs32": (4, 1, 3.0, 1), "heun2": (2, 0, 2.0, 0), "ralston2": (2, 0, 2.0, 0),
        "kutta3": (3, 0, 3.0, 0), "heun3": (3, 0, 3.0, 0), "ralston3": (3, 0, 3.0, 0), "ssprk3": (3, 0, 3.0, 0),
        "ralston4": (4, 0, 4.0, 0), "kutta4": (4, 0, 4.0, 0)}
    L = _capi.lib()
    for name, exp in expect.items():
        mid = nn.ode.method_id(name.upper())  # case-insensitive
        assert L.b200rk_method_name(mid).decode() == name
        st, fs, ad, od = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        assert L.b200rk_method_info(mid, C.byref(st), C.byref(fs), C.byref(od), C.byref(ad)) == 0
        assert (st.value, fs.value, od.value, ad.value) == exp, name
        assert (name in nn.adaptiveODE) == bool(exp[3])
    with pytest.raises(ValueError, match="rk5 is not a valid integrator"):
        nn.ode.method_id("rk5")


@pytest.mark.parametrize("method", ["dopri54", "tsit54", "vern65"])
def test_product_tableau_matches_reference_literals(method, golden_tableaux):
    """The product's own copy of the Butcher data equals the coefficients extracted from ode.nim."""
    g = golden_tableaux[method]
    c, a, b, bh = np.zeros(10), np.zeros((10, 9)), np.zeros(9), np.zeros(9)
    assert _capi.lib().b200rk_method_tableau(nn.ode.method_id(method), c.ctypes.data, a.ctypes.data, b.ctypes.data, bh.ctypes.data) == 0
    seen = 0
    for key, val in g.items():
        m = re.fullmatch(r"(c|a|b|bHat)(\d+)", key)
        kind, digits = m.group(1), m.group(2)
        if kind == "c":
            got = c[int(digits)]
        elif kind == "a":
            got = a[int(digits[0]), int(digits[1]) - 1]
        elif kind == "b":
            got = b[int(digits) - 1]
        else:
            got = bh[int(digits) - 1]
        assert float(got).hex() == float(val).hex(), (method, key)
        seen += 1
    assert seen == len(g)


def test_linspace_host_helper():  # tests/test_utils.nim:15-23
    import oracle as O
    assert nn.linspace(0.0, 10.0, 11) == [float(i) for i in range(11)]
    assert nn.linspace(-10.0, 10.0, 100) == O.linspace(-10.0, 10.0, 100).tolist()
    with pytest.raises(ValueError):
        nn.linspace(0.0, 1.0, 0)


def test_null_arguments_are_einval_not_a_crash():
    """The C-ABI never faults on a NULL handle or output pointer: EINVAL (the shim's ValueError) before any CUDA call —
    so this runs without a device, against the real library."""
    L = _capi.lib()
    null = C.c_void_p(None)
    i64 = C.c_int64(0)
    dbl = C.c_double(0.0)
    sz = C.c_size_t(0)
    calls = [
        lambda: L.b200rk_options_new(None, 1e-4, 1e-4, 1e-4, 1e-2, 1e-4, 4.0, 0.1, 0.0),
        lambda: L.b200rk_method_from_name(b"rk4", None),
        lambda: L.b200rk_method_tableau(0, None, None, None, None),
        lambda: L.b200rk_set(null, b"vec_width", 4),
        lambda: L.b200rk_get(null, b"vec_width", C.byref(i64)),
        lambda: L.b200rk_synchronize(null),
        lambda: L.b200rk_profile_reset(null),
        lambda: L.b200rk_profile_read(null, None),
        lambda: L.b200rk_ctx_stats(null, None),
        lambda: L.b200rk_vec_new(null, 8, None),
        lambda: L.b200rk_vec_fill(null, 1.0),
        lambda: L.b200rk_vec_sum(null, C.byref(dbl)),
        lambda: L.b200rk_vec_upload(null, None),
        lambda: L.b200rk_vec_download(null, None),
        lambda: L.b200rk_vec_upload_local(null, None),
        lambda: L.b200rk_vec_download_local(null, None),
        lambda: L.b200rk_vec_add(null, null, null),
        lambda: L.b200rk_vec_copy(null, null),
        lambda: L.b200rk_builtin_rhs_new(null, 0, 1.0, null, None, None),
        lambda: L.b200rk_nccl_unique_id(None),
        lambda: L.b200rk_init(None, 0),
        lambda: L.b200rk_init_distributed(None, 0, 0, 1, None),
        lambda: L.b200rk_shard_range(8, 0, 1, None, C.byref(sz)),
    ]
    for i, call in enumerate(calls):
        assert call() == _capi.EINVAL, i
    assert L.b200rk_vec_len(null) == 0 and L.b200rk_vec_local_len(null) == 0 and L.b200rk_vec_local_offset(null) == 0
    assert not L.b200rk_vec_data(null) and not L.b200rk_stream(null)
    assert L.b200rk_vec_free(null) == _capi.OK and L.b200rk_jit_rhs_free(null) == _capi.OK
    L.b200rk_destroy(null)
    L.b200rk_options_default(None)
