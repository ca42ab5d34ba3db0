"""Every script the GPU box will run parses: python files compile, shell scripts pass `bash -n`. (They can only be
executed on the B200 box; a syntax error there would cost a GPU call.)"""
import glob
import os
import py_compile
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PY = sorted(glob.glob(os.path.join(ROOT, "scripts", "*.py")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] +
            glob.glob(os.path.join(ROOT, "tests", "golden", "*.py")))
SH = sorted(glob.glob(os.path.join(ROOT, "scripts", "*.sh")))


@pytest.mark.parametrize("path", PY, ids=[os.path.relpath(p, ROOT) for p in PY])
def test_python_script_compiles(path, tmp_path):
    py_compile.compile(path, cfile=str(tmp_path / "x.pyc"), doraise=True)


@pytest.mark.parametrize("path", SH, ids=[os.path.relpath(p, ROOT) for p in SH])
def test_shell_script_parses(path):
    assert subprocess.run(["bash", "-n", path], capture_output=True, text=True).returncode == 0
