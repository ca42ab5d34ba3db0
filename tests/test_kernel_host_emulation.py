"""The product's CUDA kernels compiled by the HOST compiler and run one emulated thread at a time against the CPU
oracle (tests/host_emul/): stage accumulate, combine + error norm, the whole-attempt kernel for the built-in AND the
run-time compiled (PW_USER) right-hand sides, forward and backward in time, the RK4 kernels, and the trajectory
consumers driven by the library's own host-side plan. Same source files as the GPU build (kernels.cuh,
quad_kernels.cuh); every element-wise result must match the oracle bit for bit. It is the CPU-side gate on the kernels'
arithmetic and indexing — what remains GPU-only is concurrency (block reductions, shared-memory tiles, the cooperative
loop) and the launch plumbing. A build with FMA contraction switched on must FAIL, which shows the gate is sensitive to
exactly the property the parity claim rests on."""
import os
import re
import subprocess

import pytest

import numericalnim_b200 as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "host_emul")
LIBDIR = os.path.dirname(nn.LIB_PATH)


def _build(out, extra, src="emul_main.cpp"):
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-std=c++20", *extra, "-pthread", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable",
           f"-I{os.path.join(EMUL, 'cuda_stubs')}", f"-I{EMUL}", f"-I{os.path.join(ROOT, 'numericalnim_b200', 'csrc')}", f"-I{os.path.join(ROOT, 'oracle')}",
           os.path.join(EMUL, src), f"-L{LIBDIR}", "-lb200rk", f"-Wl,-rpath,{LIBDIR}", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def _run(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=900)
    cases = dict(re.findall(r"^case (.+) ok=(\d)$", r.stdout, flags=re.M))
    return r.returncode, cases, r.stdout + r.stderr


def test_kernels_match_the_oracle_bit_for_bit(tmp_path):
    rc, cases, out = _run(_build(str(tmp_path / "emul_main"), ["-O1", "-ffp-contract=off"]))
    assert len(cases) >= 30, out
    bad = [k for k, v in cases.items() if v != "1"]
    assert rc == 0 and not bad, bad
    for family in ("stage_kernel", "finish_kernel", "finish_pf_kernel", "fused_attempt", "source rhs", "user_rk4_kernel", "cumtrapz_kernel", "hermite_many_kernel"):
        assert any(family in k for k in cases), family
    exe = str(tmp_path / "emul_main")
    for seed in range(1, 13):  # the same cases on other random inputs
        rc, cases, out = _run(exe, str(seed))
        assert rc == 0 and all(v == "1" for v in cases.values()), (seed, [k for k, v in cases.items() if v != "1"])


def test_the_gate_detects_fma_contraction(tmp_path):
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU has no FMA instructions to contract into")
    rc, cases, _ = _run(_build(str(tmp_path / "emul_fma"), ["-O2", "-march=native", "-ffp-contract=fast"]))
    bad = [k for k, v in cases.items() if v != "1"]
    assert rc != 0 and len(bad) >= 10, (rc, bad)


# ---------------------------------------------------------------------------------------------------
# threaded emulation: one host thread per CUDA thread, real barriers / shuffles / atomics (emul.hpp, EMUL_MT)
# ---------------------------------------------------------------------------------------------------
def test_interacting_kernels_under_threaded_emulation(tmp_path):
    """The kernels whose threads interact — the two-stage reduction, the shared-memory Lorenz-96 stage kernel, and the
    device-resident driver loop as a one-block cooperative grid (step / attempt / rejection / limiter counts equal to
    the oracle's ODESolver) — plus all the cases above once more with the REAL reduction instead of a running sum."""
    rc, cases, out = _run(_build(str(tmp_path / "emul_mt"), ["-O1", "-ffp-contract=off"], "emul_mt_main.cpp"))
    assert rc == 0 and len(cases) >= 8 and all(v == "1" for v in cases.values()), out
    for family in ("neq_count_kernel", "stage_l96_kernel", "l96_attempt_kernel", "fused_run_kernel", "source rhs"):
        assert any(family in k for k in cases), family
    rc, cases, out = _run(_build(str(tmp_path / "emul_main_mt"), ["-O1", "-ffp-contract=off", "-DEMUL_MT"]))
    assert rc == 0 and len(cases) >= 30 and all(v == "1" for v in cases.values()), out


def test_thread_sanitizer_finds_no_race_in_the_kernels(tmp_path):
    """Race detection without a GPU: the threaded emulation under ThreadSanitizer. The kernels' shared- and
    global-memory traffic must be race-free; a tile kernel with its __syncthreads removed must be flagged."""
    probe = tmp_path / "probe.cpp"
    probe.write_text("int main() { return 0; }\n")
    if subprocess.run(["g++", "-fsanitize=thread", str(probe), "-o", str(tmp_path / "probe")], capture_output=True).returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available")
    exe = _build(str(tmp_path / "emul_mt_tsan"), ["-O1", "-g", "-ffp-contract=off", "-fsanitize=thread", "-Wno-tsan"], "emul_mt_main.cpp")
    rc, cases, out = _run(exe)
    if "FATAL: ThreadSanitizer" in out:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + out.strip().splitlines()[0])
    assert rc == 0 and len(cases) >= 8 and all(v == "1" for v in cases.values()), out[-3000:]
    assert "WARNING: ThreadSanitizer" not in out, out[-3000:]
    _, _, racy = _run(exe, "racy")
    assert "WARNING: ThreadSanitizer: data race" in racy, "the positive control was not flagged"
