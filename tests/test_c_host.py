"""The C-ABI from plain C (what Nim's importc sees): the header is valid C99, the demo host program compiles
and links against libb200rk.so on the CPU box, and on the GPU box it runs and agrees with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "numericalnim_b200", "lib")
SRC = os.path.join(ROOT, "examples", "c_host_demo.c")


def _build(tmp_path):
    exe = str(tmp_path / "c_host_demo")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O2", f"-I{INC}", SRC, f"-L{LIBDIR}", "-lb200rk", "-lm",
                    f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True, capture_output=True, text=True)
    return exe


def test_header_is_plain_c99(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "b200rk.h"\nint main(void) { b200rk_options o; (void)o; return B200RK_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", f"-I{INC}", str(src)], check=True)


def test_c_host_program_links(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", ["dopri54", "tsit54", "vern65", "rk4"])
def test_c_host_program_matches_oracle(tmp_path, integrator):
    import oracle as O
    exe = _build(tmp_path)
    n = 4096
    r = subprocess.run([exe, integrator, str(n)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    head = dict(kv.split("=") for kv in r.stdout.splitlines()[0].split())
    lam = 0.1 + 9.9 * np.arange(n) / (n - 1)
    import math
    y0 = np.array([1.0 + 0.5 * math.sin(2.0 * 3.14159265358979323846 * float(i) / float(n)) for i in range(n)])  # libm sin, as in the C program
    ref = O.solve_vector(integrator, O.rhs_diag_linear(lam), y0, [0.0, 2.0], O.new_options(dt=1e-2, absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8))
    assert int(head["steps"]) == ref.stats.steps and int(head["rejected"]) == ref.stats.rejected and int(head["n_out"]) == 2
    assert int(head["launches"]) > 0
    for m in re.finditer(r"y\[(\d+)\]=(\S+)", r.stdout):
        i, v = int(m.group(1)), float.fromhex(m.group(2))
        if integrator == "rk4":
            assert v == ref.y[-1][i]  # fixed step: bit-identical
        else:
            assert abs(v - ref.y[-1][i]) <= 1e-9 * abs(ref.y[-1][i]) + 1e-13
