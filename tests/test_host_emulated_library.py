"""The GPU parity suites, run on the CPU against the HOST-EMULATED library (tests/host_emul/build_emul_lib.py): the
product's own host sources (drivers, planners, launch code) and kernels compiled by g++ — the only source change is the
`<<<grid, threads>>>` launch syntax, which becomes a loop over emulated threads (a real thread team with barriers for
the shared-memory Lorenz-96 kernel and the cooperative device loop); CUDA runtime calls resolve to a fake runtime whose
"device" memory is host memory. It is what lets a change to host logic or kernel arithmetic be checked end to end
against the oracle before any GPU time is spent — and what verified, this round, everything written after the GPU budget
ran out.

TEST INFRASTRUCTURE ONLY: the library is built under tests/, selected by a pytest option, and the product package has no
way to load it (numericalnim_b200 still raises without a CUDA device; tests/test_capi_host.py checks that). What it
cannot cover: NVRTC / module loading (right-hand sides from source), multi-CTA concurrency, peer mailboxes, performance."""
import importlib.util
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITES = ["tests/test_gpu_parity.py", "tests/test_gpu_fuzz.py", "tests/test_gpu_quadrature.py", "tests/test_gpu_device_loop.py",
          "tests/test_gpu_fused_paths.py"]


def test_gpu_suites_pass_under_host_emulation():
    cmd = [sys.executable, "-m", "pytest", *SUITES, "-q", "-m", "gpu", "--host-emulation", "-p", "no:cacheprovider"]
    if importlib.util.find_spec("xdist") is not None:
        cmd += ["-n", str(min(6, os.cpu_count() or 1))]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=3000, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-2000:]
    counts = {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|error|errors|xpassed|xfailed|skipped)", tail)}
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert counts.get("failed", 0) == 0 and counts.get("error", 0) == 0 and counts.get("errors", 0) == 0, tail
    assert counts.get("passed", 0) >= 225, tail            # parity 147 + fuzz 3 + quadrature 40 + device loop 9 + fused paths 26
    assert counts.get("xpassed", 0) + counts.get("xfailed", 0) == 0, tail


def test_emulated_library_is_not_reachable_from_the_product():
    """No CPU fallback: the package names neither the emulated library nor the switch that selects it."""
    pkg = os.path.join(ROOT, "numericalnim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".nim")) and "lib" + os.sep + "obj" not in dirpath:
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "b200rk_emul" not in src and "host_emul" not in src.replace("tests/host_emul", "") and "--host-emulation" not in src, f


def test_compiled_language_hosts_run_against_the_emulated_library(tmp_path):
    """The C and C++ demos (what a Nim / C / C++ host does with the C-ABI), linked against the emulated library and run
    on the CPU: the reference's Vector ODE tests through C++ closures, the plain-C solve, and the trajectory consumers
    through the C++ mirror (the demo whose first GPU run is still pending). The demos that hand a right-hand side over
    as source need NVRTC + a GPU and stay GPU-only."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "host_emul"))
    import build_emul_lib
    lib = build_emul_lib.build()
    libdir, inc = os.path.dirname(lib), os.path.join(ROOT, "include")
    link = [f"-L{libdir}", "-lb200rk_emul", f"-Wl,-rpath,{libdir}"]

    def run(src, compiler, std, args=(), extra=()):
        exe = str(tmp_path / os.path.basename(src).split(".")[0])
        subprocess.run([compiler, std, "-O2", f"-I{inc}", os.path.join(ROOT, "examples", src), *link, *extra, "-o", exe], check=True, capture_output=True, text=True)
        r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        return r.stdout

    assert "quadrature trajectory_ok=1 function_variant_ok=1" in run("cpp_quadrature_demo.cpp", "g++", "-std=c++17")
    out = run("cpp_host_demo.cpp", "g++", "-std=c++17")
    assert out.count(" ok=1 ") == 8 and "errors raised=4 of 4" in out
    out = run("c_host_demo.c", "gcc", "-std=c99", args=("tsit54", "4096"), extra=("-lm",))
    assert "n_out=2" in out and "max_abs_err_vs_exact" in out


def _gcc_file(name):
    r = subprocess.run(["gcc", "-print-file-name=" + name], capture_output=True, text=True)
    p = r.stdout.strip()
    return p if r.returncode == 0 and os.path.isabs(p) and os.path.exists(p) else None


def test_memory_errors_under_address_sanitizer(tmp_path):
    """memcheck without a GPU: the emulated library built with AddressSanitizer. "Device" memory is host memory, so an
    out-of-bounds access by a kernel (ragged sizes, tile seams, pointer tables) or by the host code is reported the way
    compute-sanitizer would. Default: the trajectory-consumer, experimental and fuzz suites (seconds);
    B200RK_TEST_ASAN=full runs every emulated GPU suite (207 cases, ~4 min — clean at the end of round 1)."""
    import pytest
    asan, stdcpp = _gcc_file("libasan.so"), _gcc_file("libstdc++.so.6")
    if not asan or not stdcpp:
        pytest.skip("AddressSanitizer runtime not available")
    suites = SUITES if os.environ.get("B200RK_TEST_ASAN") == "full" else ["tests/test_gpu_quadrature.py", "tests/test_gpu_fused_paths.py", "tests/test_gpu_fuzz.py"]
    log = str(tmp_path / "asan")
    env = dict(os.environ, LD_PRELOAD=f"{asan} {stdcpp}", ASAN_OPTIONS=f"detect_leaks=0:log_path={log}", B200RK_TEST_EMULATION_SANITIZE="address")
    cmd = [sys.executable, "-m", "pytest", *suites, "-q", "-m", "gpu", "--host-emulation", "-p", "no:cacheprovider"]
    if importlib.util.find_spec("xdist") is not None:
        cmd += ["-n", str(min(6, os.cpu_count() or 1))]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=3000, cwd=ROOT, env=env)
    reports = [f for f in os.listdir(tmp_path) if f.startswith("asan")]
    detail = "".join(open(os.path.join(tmp_path, f)).read()[:3000] for f in reports[:2])
    if "Shadow memory range interleaves" in detail or "ASan runtime does not come first" in r.stderr + detail:
        pytest.skip("AddressSanitizer cannot be preloaded into this interpreter")
    assert not reports, detail
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-2000:]
    assert r.returncode == 0 and " failed" not in tail and " error" not in tail, r.stdout[-3000:] + r.stderr[-2000:]
