"""The compiled-language host mirror (include/numericalnim_b200.hpp): compiles and links on the CPU box; on the
GPU box it runs the reference's Vector ODE tests (tests/test_ode.nim:139-197) written as C++ closures and the
results agree with the CPU oracle."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "numericalnim_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "cpp_host_demo")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O2", f"-I{INC}", os.path.join(ROOT, "examples", "cpp_host_demo.cpp"),
                    f"-L{LIBDIR}", "-lb200rk", f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True, capture_output=True, text=True)
    return exe


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_cpp_host_mirror_runs_reference_vector_tests(tmp_path):
    import oracle as O
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    cases = [dict(kv.split("=") for kv in line.split()[1:]) for line in r.stdout.splitlines() if line.startswith("case ")]
    assert len(cases) == 8 and all(c["ok"] == "1" for c in cases)
    assert "errors raised=4 of 4" in r.stdout
    ts = O.linspace(-10.0, 10.0, 100)
    for c in cases:
        opts = O.new_options(relTol=1e-8, dt=1e-2) if c["options"] == "ooVector" else O.new_options()
        ref = O.solve_vector(c["integrator"], O.rhs_scale(-0.1), [1.0, 1.0, 1.0], ts, opts)
        assert int(c["steps"]) == ref.stats.steps and int(c["rhs_evals"]) == ref.stats.rhs_evals, c
        y10 = float.fromhex(c["y10"])
        if c["integrator"] in ("rk4", "heun2"):
            assert y10 == ref.y[-1][0], c  # fixed step: bit-identical through user closures too
        else:
            assert abs(y10 - ref.y[-1][0]) <= 1e-9 * abs(ref.y[-1][0]), c
