"""The run-time compiled right-hand sides and the trajectory consumers through the compiled-language mirror
(examples/cpp_jit_demo.cpp, examples/cpp_quadrature_demo.cpp): both compile and link on the CPU box; on the GPU box
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
