"""Properties of the SHIPPED machine code, read from the build objects with cuobjdump (no GPU needed):

* no FMA contraction anywhere: DFMA appears only in kernels that contain an IEEE division / square root / pow
  expansion (MUFU.RCP64H / RSQ64H seeds refined with FMAs — that is how __ddiv_rn is implemented), never in the
  division-free kernels (stage accumulate, RK4 combine, Hermite, the Vector operators, the trajectory scans). This is
  the binary-level statement of "same multiply/add sequence as the reference's CPU arithmetic";
* the streaming kernels use the 256-bit global accesses (LDG.E...256 / STG.E...256) the design relies on;
* the hot kernels do not spill (no local-memory stack), except the device-resident loop (a few dozen bytes)."""
import collections
import glob
import os
import re
import shutil
import subprocess

import pytest

import numericalnim_b200 as nn

OBJ = os.path.join(os.path.dirname(nn.LIB_PATH), "obj")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not glob.glob(os.path.join(OBJ, "*.o")),
                                reason="cuobjdump or the build objects are not available")


def _device_objects():
    """Build objects that carry device code (kernels_src.o / stencil_src.o are the embedded header texts: host data only)."""
    return [p for p in sorted(glob.glob(os.path.join(OBJ, "*.o"))) if os.path.basename(p) not in ("kernels_src.o", "stencil_src.o")]


def _kernels():
    out = {}
    for path in _device_objects():
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        fn = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = m.group(1)
                out[fn] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and fn:
                out[fn][m.group(1)] += 1
    return out


@pytest.fixture(scope="module")
def kernels():
    k = _kernels()
    assert len(k) > 150
    return k


def _count(c, prefix):
    return sum(v for k, v in c.items() if k.startswith(prefix))


def test_no_fma_outside_division_expansions(kernels):
    offenders = [fn for fn, c in kernels.items() if _count(c, "DFMA") and not _count(c, "MUFU")]
    assert not offenders, offenders
    # the division-free kernel families contain no DFMA at all, and do contain separate DMUL and DADD
    for family in ("stage_kernel", "rk4_final_kernel", "hermite_kernel", "hermite_many_kernel", "cumtrapz_kernel", "simpson_scan_kernel",
                   "trapz_step_kernel", "lorenz96_kernel", "stage_l96_kernel", "fused_rk4_kernel"):
        members = {fn: c for fn, c in kernels.items() if family in fn}
        assert members, family
        for fn, c in members.items():
            assert _count(c, "DFMA") == 0, fn
            assert _count(c, "DMUL") > 0 and _count(c, "DADD") > 0, fn


def test_streaming_kernels_use_256_bit_accesses(kernels):
    wide = {"stage_kernelILi5ELi4E": 1, "finish_kernelILi6ELi4E": 0, "fused_attempt_kernelILi0ELi1ELi4E": 1, "fused_run_kernelILi0ELi1ELi4E": 1,
            "cumtrapz_kernelILi4E": 1, "simpson_scan_kernelILi4E": 1, "hermite_many_kernelILi4E": 1}
    for tag, stores in wide.items():
        members = [c for fn, c in kernels.items() if tag in fn]
        assert members, tag
        for c in members:
            assert any(k.startswith("LDG") and ".256" in k for k in c), tag
            if stores:
                assert any(k.startswith("STG") and ".256" in k for k in c), tag


def test_hot_kernels_do_not_spill():
    res = ""
    for path in _device_objects():
        res += subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True, check=True).stdout
    usage = {m[0]: (int(m[1]), int(m[2])) for m in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", res)}
    assert len(usage) > 150
    for fn, (reg, stack) in usage.items():
        if "fused_run_kernel" in fn:
            strict = "fused_run_kernelILi1E" in fn or "fused_run_kernelILi4E" in fn   # strict_zeros variants (all tableau terms kept)
            assert reg <= 128, (fn, reg, stack)                                       # __launch_bounds__(256, 2): two CTAs per SM must fit
            assert stack <= (160 if strict else 64), (fn, reg, stack)                 # default patterns: at most a few spilled doubles
        elif any(t in fn for t in ("stage_kernel", "finish_kernel", "fused_attempt_kernel", "ewise_kernel", "cumtrapz_kernel", "simpson_scan_kernel",
                                    "hermite_many_kernel", "stage_l96_kernel", "rk4_final_kernel", "hermite_kernel")):
            if "ewise_kernelILi8E" in fn or "ewise_kernelILi3E" in fn:   # `/`: the division's slow path may keep two doubles on the stack
                assert stack <= 16, (fn, reg, stack)
                continue
            assert stack == 0, (fn, reg, stack)
