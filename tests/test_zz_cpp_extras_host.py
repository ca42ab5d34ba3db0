"""The run-time compiled right-hand sides and the trajectory consumers through the compiled-language mirror and from
plain C (examples/cpp_jit_demo.cpp, examples/cpp_quadrature_demo.cpp, examples/c_extras_demo.c): both compile and link on the CPU box; on the GPU box
they run and check themselves against the closure path / the reference's tolerances (tests/test_integrate.nim:67-95).
(Named to run after the kernel-level parity suites.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "numericalnim_b200", "lib")


def _build(tmp_path, name):
    exe = str(tmp_path / name)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O2", f"-I{INC}", os.path.join(ROOT, "examples", name + ".cpp"),
                    f"-L{LIBDIR}", "-lb200rk", f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True, capture_output=True, text=True)
    return exe


@pytest.mark.parametrize("name", ["cpp_jit_demo", "cpp_quadrature_demo"])
def test_cpp_demo_compiles_and_links(tmp_path, name):
    assert os.path.exists(_build(tmp_path, name))


@pytest.mark.gpu
def test_cpp_jit_demo_runs(tmp_path):
    r = subprocess.run([_build(tmp_path, "cpp_jit_demo")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    jit = [dict(kv.split("=") for kv in line.split()[1:]) for line in r.stdout.splitlines() if line.startswith("jit ")]
    assert len(jit) == 1 and jit[0]["same_bits_as_closure"] == "1" and jit[0]["dopri54_ok"] == "1" and jit[0]["bad_expression_raises"] == "1"
    assert int(jit[0]["launches_jit"]) < int(jit[0]["launches_closure"])  # fused RK4 step: 1 launch instead of 8


@pytest.mark.gpu
def test_cpp_quadrature_demo_runs(tmp_path):
    r = subprocess.run([_build(tmp_path, "cpp_quadrature_demo")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "quadrature trajectory_ok=1 function_variant_ok=1" in r.stdout


# ---- the same entry points from plain C99 (what the Nim {.importc, cdecl.} shim binds) ----------------------------------
def _build_c(tmp_path):
    exe = str(tmp_path / "c_extras_demo")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O2", f"-I{INC}", os.path.join(ROOT, "examples", "c_extras_demo.c"),
                    f"-L{LIBDIR}", "-lb200rk", "-lm", f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True, capture_output=True, text=True)
    return exe


def test_c_extras_demo_compiles_and_links(tmp_path):
    assert os.path.exists(_build_c(tmp_path))


@pytest.mark.gpu
def test_c_extras_demo_runs(tmp_path):
    """Logistic growth with the right-hand side given as source, its cumulative Simpson integral and an interpolated
    state, each against the closed form (tolerances set from the same computation done with the CPU oracle:
    6e-13 / 3e-7 / 2e-8 observed there)."""
    r = subprocess.run([_build_c(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad_expression_rc=1 is_einval=1" in r.stdout
    line = [l for l in r.stdout.splitlines() if l.startswith("steps=")][0]
    vals = dict(kv.split("=") for kv in line.split())
    assert float(vals["err_solution"]) < 1e-6 and float(vals["err_integral"]) < 1e-5 and float(vals["err_interpolated"]) < 1e-5
    # the heat equation on a ring given as a stencil expression: one kernel per attempt, exact decay of a Fourier mode
    assert "bad_stencil_rc=1 is_einval=1" in r.stdout
    sv = dict(kv.split("=") for kv in [l for l in r.stdout.splitlines() if l.startswith("stencil_attempts=")][0].split())
    assert float(sv["err_stencil"]) < 1e-7 and int(sv["stencil_launches"]) <= int(sv["stencil_attempts"]) + 8
