"""world = 2 on the CPU: two rank threads of one process drive the HOST-EMULATED library (tests/host_emul/) through an
in-process stand-in for NCCL (tests/host_emul/fake_nccl.cpp) — the host code no single-rank run reaches: sharding, the
ncclAllReduce form of the error norm, the Lorenz-96 ring halo per right-hand-side evaluation, and the one-kernel
Lorenz-96 attempt / RK4 step with one halo exchange per call, every shard against the unsharded oracle (bit for bit
where the step size is given). TEST INFRASTRUCTURE ONLY; CUDA IPC is not emulated, so the peer-mailbox all-reduce and the
peer-mapped halo remain GPU-only (scripts/multi_gpu_check.py)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_against_the_unsharded_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_emul", "two_rank_emul.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    cases = dict(re.findall(r"^case (.+) ok=(\d)$", r.stdout, flags=re.M))
    assert r.returncode == 0 and len(cases) >= 16 and all(v == "1" for v in cases.values()), r.stdout[-3000:] + r.stderr[-2000:]
    for family in ("diag dopri54 sharded", "3-element halo per evaluation", "one halo exchange per call", "one-kernel step bitwise", "rk4 one-kernel step"):
        assert any(family in k for k in cases), family
    # one grouped exchange per step instead of one per right-hand-side evaluation
    info = dict(re.findall(r"^info lorenz96 n=1000 fuse_stencil_attempt=(\d): .*collectives=(\d+)$", r.stdout, flags=re.M))
    assert int(info["1"]) < int(info["0"]) / 2, info
