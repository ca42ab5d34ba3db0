"""The trajectory consumers through the compiled-language mirror (examples/cpp_quadrature_demo.cpp): compiles and links
on the CPU box; on the GPU box it integrates a solved trajectory and a closure integrand and checks the reference's
tolerances (tests/test_integrate.nim:67-95). (Named to run after the kernel-level parity suites.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "numericalnim_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "cpp_quadrature_demo")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O2", f"-I{INC}", os.path.join(ROOT, "examples", "cpp_quadrature_demo.cpp"),
                    f"-L{LIBDIR}", "-lb200rk", f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True, capture_output=True, text=True)
    return exe


def test_cpp_quadrature_demo_compiles_and_links(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_cpp_quadrature_demo_runs(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "quadrature trajectory_ok=1 function_variant_ok=1" in r.stdout
