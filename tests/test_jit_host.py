"""Host-side checks of the run-time compiled right-hand sides (numericalnim_b200/csrc/jit.cu). NVRTC needs no GPU:
every translation unit the library would load on the B200 is compiled here for sm_100a, and its SASS is inspected
(no FMA contraction of the user's a*b+c; 256-bit loads in the fused kernels). No compute calls."""
import re
import shutil
import subprocess

import pytest

import numericalnim_b200 as nn

PATTERNS = {0: "dopri54", 1: "dopri54 strict", 2: "tsit54", 3: "vern65", 4: "vern65 strict"}


def _sass(cubin: bytes, tmp_path, name="u.cubin") -> str:
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    p = tmp_path / name
    p.write_bytes(cubin)
    return subprocess.run(["cuobjdump", "-sass", str(p)], capture_output=True, text=True, check=True).stdout


def _per_function(sass: str) -> dict:
    out, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            out[fn] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            out[fn].append(m.group(1))
    return out


def test_base_unit_compiles_and_names_its_kernels():
    cubin, log = nn.jitCompileOnly("c0*y*(1.0 - y/p0) + c1*t", n_vec=1, n_scalar=2, pattern=-1)
    assert len(cubin) > 10_000 and cubin[:4] == b"\x7fELF"
    names = [l.split()[1] for l in log.splitlines() if l.startswith("kernel ")]
    assert len(names) == 3
    assert sum("user_rhs_kernel" in n for n in names) == 2 and sum("user_rk4_kernel" in n for n in names) == 1


@pytest.mark.parametrize("pattern", sorted(PATTERNS))
def test_fused_units_compile_for_every_pattern(pattern):
    """The whole-attempt kernel and the device-resident loop, instantiated around a user expression, for each pair."""
    cubin, log = nn.jitCompileOnly("-(p0*y) + p1*c0", n_vec=2, n_scalar=1, pattern=pattern)
    names = [l.split()[1] for l in log.splitlines() if l.startswith("kernel ")]
    assert len(names) == 4, log
    assert sum("fused_attempt_kernel" in n for n in names) == 2 and sum("fused_run_kernel" in n for n in names) == 2
    assert all(f"ILi{pattern}ELi2E" in n for n in names), names  # <PAT, PW_USER, ...>


def test_no_parameters_and_time_only_expressions_compile():
    nn.jitCompileOnly("-0.1*y", 0, 0, -1)
    nn.jitCompileOnly("cos(t)", 0, 0, 0)
    nn.jitCompileOnly("c0*y + c7", 0, 8, -1)
    nn.jitCompileOnly("p0 + p1 + p2 + p3", 4, 0, 2)


def test_wrong_expression_is_a_value_error_with_the_compiler_log():
    with pytest.raises(ValueError, match=r'identifier "yy" is undefined'):
        nn.jitCompileOnly("c0*yy", 0, 1, -1)
    with pytest.raises(ValueError, match="p1"):  # only p0 exists
        nn.jitCompileOnly("p0*y + p1", 1, 0, -1)
    for bad in ("", "y; y", "y } {", "#include <x>\ny"):
        with pytest.raises(ValueError):
            nn.jitCompileOnly(bad, 0, 0, -1)
    with pytest.raises(ValueError, match="at most 4"):
        nn.jitCompileOnly("y", 5, 0, -1)
    with pytest.raises(ValueError, match="at most 8"):
        nn.jitCompileOnly("y", 0, 9, -1)
    with pytest.raises(ValueError, match="pattern"):
        nn.jitCompileOnly("y", 0, 0, 7)


def test_user_multiply_add_is_not_contracted(tmp_path):
    """--fmad=false: `p0*y + c0` must stay DMUL + DADD (bit parity with the reference's CPU arithmetic)."""
    cubin, _ = nn.jitCompileOnly("p0*y + c0", 1, 1, -1)
    fns = _per_function(_sass(cubin, tmp_path))
    rhs = [ops for name, ops in fns.items() if "user_rhs_kernel" in name]
    assert len(rhs) == 2
    for ops in rhs:
        assert not any(o.startswith("DFMA") for o in ops), "user expression was contracted into FMA"
        assert any(o.startswith("DMUL") for o in ops) and any(o.startswith("DADD") for o in ops)
    wide = [ops for name, ops in fns.items() if "user_rhs_kernelILi4E" in name][0]
    assert any(".256" in o and o.startswith("LDG") for o in wide) and any(".256" in o and o.startswith("STG") for o in wide)


def test_jit_fused_kernel_matches_the_builtin_instruction_mix(tmp_path):
    """`-(p0*y)` compiled at run time is the built-in diag-linear right-hand side: the whole-attempt kernel NVRTC
    produces has the same fp64 instruction mix as the one nvcc built into the library."""
    import os
    obj = os.path.join(os.path.dirname(nn.LIB_PATH), "obj", "executor.o")
    if not os.path.exists(obj) or shutil.which("cuobjdump") is None:
        pytest.skip("build objects / cuobjdump not available")
    cubin, _ = nn.jitCompileOnly("-(p0*y)", 1, 0, 0)
    jit = _per_function(_sass(cubin, tmp_path))
    ref = _per_function(subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout)

    def mix(ops):
        keys = ("DMUL", "DADD", "DFMA", "MUFU", "LDG.E.NA", "STG.E.NA")
        return {k: sum(o.startswith(k) for o in ops) for k in keys}

    j = [ops for n, ops in jit.items() if "fused_attempt_kernelILi0ELi2ELi4E" in n][0]
    r = [ops for n, ops in ref.items() if "fused_attempt_kernelILi0ELi1ELi4ELi256ELi0E" in n][0]
    assert mix(j) == mix(r), (mix(j), mix(r))


def test_compiles_with_an_older_nvrtc_already_in_the_process(monkeypatch):
    """PyTorch bundles the libnvrtc.so.12 of its own (older) toolkit; once it is loaded, a dlopen by SONAME would
    return that copy, whose ptxas rejects the kernels' 256-bit accesses. jit.cu opens the toolkit's NVRTC by path."""
    torch = pytest.importorskip("torch")
    try:
        import ctypes
        import glob
        import os
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_nvrtc", "lib", "libnvrtc.so.12"))
        if cands:
            ctypes.CDLL(cands[0], mode=ctypes.RTLD_GLOBAL)
    except Exception:  # noqa: BLE001 — best effort: the point is only to have another NVRTC mapped
        pass
    monkeypatch.setenv("B200RK_JIT_CACHE", "off")
    cubin, log = nn.jitCompileOnly("c0*y*(1.0 - y/p0) + 0.125", 1, 1, 0)  # an expression no other test compiles
    assert cubin[:4] == b"\x7fELF" and log.count("kernel _Z") == 4


def test_kernel_argument_blocks_have_the_same_size_in_both_compilations(tmp_path):
    """The host fills FusedArgs / RunArgs as laid out by nvcc's host compiler and passes them by value to kernels
    NVRTC compiled: the parameter-block sizes recorded in the two ELF images must agree."""
    import os
    obj = os.path.join(os.path.dirname(nn.LIB_PATH), "obj", "executor.o")
    if not os.path.exists(obj) or shutil.which("cuobjdump") is None:
        pytest.skip("build objects / cuobjdump not available")

    def param_sizes(path):
        out = subprocess.run(["cuobjdump", "-elf", path], capture_output=True, text=True, check=True).stdout
        sizes, cur, want = {}, None, False
        for line in out.splitlines():
            m = re.match(r"\s*\.nv\.info\.(\S+)", line)
            if m:
                cur = m.group(1)
            if "EIATTR_PARAM_CBANK" in line:
                want = True
                continue
            if want and "Value:" in line and cur:
                sizes[cur] = int(line.split()[-1], 16) >> 16  # high half of the second word = size in bytes
                want = False
        return sizes

    ref = param_sizes(obj)
    for pat in (0, 2, 3):
        cubin, _ = nn.jitCompileOnly("-(p0*y)", 1, 0, pat)
        p = tmp_path / f"u{pat}.cubin"
        p.write_bytes(cubin)
        for name, size in param_sizes(str(p)).items():
            twin = name.replace(f"ILi{pat}ELi2E", f"ILi{pat}ELi1E")  # PW_USER -> PW_DIAG instance built by nvcc
            assert twin in ref, twin
            assert ref[twin] == size and size > 800, (name, size, ref[twin])


# ---- stencil right-hand sides from source (b200rk_jit_stencil_rhs_new): the same host-only compile ----------------------
STENCILS = [("((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, 0, 1),                    # Lorenz-96
            ("c0*((Y(-1) - 2.0*Y(0)) + Y(1))*p0 + c1*t", 1, 1, 1, 2),                # diffusion with a coefficient field and a source term
            ("-(c0*(Y(0) - Y(-1)))", 1, 0, 0, 1),                                     # upwind advection: one-sided
            ("(Y(-3) + Y(2)) - 2.0*Y(0)", 3, 2, 0, 0),
            ("c0*(Y(3) - Y(0))", 0, 3, 0, 1), ("((Y(-8) + Y(8)) - 2.0*Y(0))*c0", 8, 8, 0, 1), ("c0*Y(0) + c1*t", 0, 0, 0, 2)]   # right-only, maximum radii, element-local


@pytest.mark.parametrize("expr,rl,rr,nv,ns", STENCILS)
def test_stencil_units_compile_for_the_rhs_kernel_and_every_pair(expr, rl, rr, nv, ns):
    n, log = nn.jitStencilCompileOnly(expr, rl, rr, nv, ns, -1)
    assert n > 1000 and "user_stencil_rhs_kernel" in log
    for pattern in (0, 2, 3):
        n, log = nn.jitStencilCompileOnly(expr, rl, rr, nv, ns, pattern)
        assert n > 1000 and "ustencil_attempt_kernel" in log, log


def test_stencil_offset_outside_the_declared_radii_is_a_value_error():
    with pytest.raises(ValueError) as e:
        nn.jitStencilCompileOnly("Y(2) - Y(0)", 1, 1, 0, 0, -1)
    assert "Y(2)" in str(e.value)
    with pytest.raises(ValueError):
        nn.jitStencilCompileOnly("Y(0)", 9, 0, 0, 0, -1)     # radii are limited to 0..8
    with pytest.raises(ValueError):
        nn.jitStencilCompileOnly("Y(0); y", 1, 1, 0, 0, -1)  # a single expression
