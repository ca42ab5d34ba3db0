"""world = 2 on the CPU: two rank threads of one process drive the HOST-EMULATED library (tests/host_emul/) through an
in-process stand-in for NCCL (tests/host_emul/fake_nccl.cpp) — the host code no single-rank run reaches: sharding, both
forms of the error-norm all-reduce (peer mailboxes mapped through an emulated CUDA IPC where a handle is the pointer, and
ncclAllReduce with B200RK_P2P=0), the Lorenz-96 ring halo per right-hand-side evaluation, and the one-kernel Lorenz-96
attempt / RK4 step with its halo exchanged once per call or read in place from the peer-mapped neighbour (peer_view_open,
stream barriers, collective close) — every shard against the unsharded oracle, bit for bit where the step size is given.
TEST INFRASTRUCTURE ONLY: what it cannot show is NVLink / CUDA-IPC behaviour itself and cross-GPU memory ordering
(scripts/multi_gpu_check.py on the GPU box)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,p2p", [(2, 1), (2, 0), (3, 1), (4, 0)])   # 3+: the ring neighbours are two different peers
def test_rank_threads_against_the_unsharded_oracle(world, p2p):
    env = dict(os.environ, B200RK_TEST_EMUL_WORLD=str(world))
    env.pop("B200RK_P2P", None)
    if not p2p:
        env["B200RK_P2P"] = "0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_emul", "multi_rank_emul.py")], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    cases = dict(re.findall(r"^case (.+) ok=(\d)$", r.stdout, flags=re.M))
    assert r.returncode == 0 and len(cases) >= 20 and all(v == "1" for v in cases.values()), r.stdout[-3000:] + r.stderr[-2000:]
    assert f"info p2p={p2p}" in r.stdout
    families = ["diag dopri54 sharded", "sum(v) sharded", "cumsimpson sharded bitwise", "impure duplicate", "3-element halo per evaluation", "one halo exchange per call", "one-kernel step bitwise", "rk4 one-kernel step"]
    if p2p:
        families.append("halo read in place from the peer-mapped neighbour")
    for family in families:
        assert any(family in k for k in cases), family
    # collectives per solve (n = 1000, 9 attempts): a 3-element halo in front of every right-hand-side evaluation > one
    # grouped exchange per step > none at all on the stencil path (what remains: the error-norm all-reduce per attempt,
    # opening / closing the mapping, one barrier)
    info = {(int(a), int(b)): int(c) for a, b, c in re.findall(r"^info lorenz96 n=1000 fuse_stencil_attempt=(\d) peer_halo=(\d): .*collectives=(\d+)$", r.stdout, flags=re.M)}
    assert info[(1, 0)] < info[(0, 1)] / 2, info
    if p2p:
        assert info[(1, 1)] < info[(1, 0)], info


def test_three_ranks_under_address_sanitizer(tmp_path):
    """The same run on the AddressSanitizer build of the emulated library: the halo staging buffer, the handle tables of
    peer_view_open and the shard-edge indexing are host code no single-rank run touches."""
    def gcc_file(name):
        p = subprocess.run(["gcc", "-print-file-name=" + name], capture_output=True, text=True).stdout.strip()
        return p if os.path.isabs(p) and os.path.exists(p) else ""
    asan, stdcpp = gcc_file("libasan.so"), gcc_file("libstdc++.so.6")
    if not asan or not stdcpp:
        pytest.skip("AddressSanitizer runtime not available")
    log = str(tmp_path / "asan")
    env = dict(os.environ, LD_PRELOAD=f"{asan} {stdcpp}", ASAN_OPTIONS=f"detect_leaks=0:log_path={log}", B200RK_TEST_EMULATION_SANITIZE="address",
               B200RK_TEST_EMUL_WORLD="3")
    env.pop("B200RK_P2P", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_emul", "multi_rank_emul.py")], capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    reports = [f for f in os.listdir(tmp_path) if f.startswith("asan")]
    detail = "".join(open(os.path.join(tmp_path, f)).read()[:3000] for f in reports[:2])
    if "Shadow memory range interleaves" in detail or "ASan runtime does not come first" in r.stderr + detail:
        pytest.skip("AddressSanitizer cannot be preloaded into this interpreter")
    assert not reports, detail
    assert r.returncode == 0 and "failures=0 hung=0" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
