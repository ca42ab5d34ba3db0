#!/bin/bash
# Round 2, GPU call: stencil right-hand sides from source — first GPU run of the parity tests + throughput next to the built-in Lorenz-96.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
timeout 900 python -m pytest tests/test_gpu_stencil_from_source.py -q -x -p no:cacheprovider 2>&1 | tail -15 | cut -c1-400
timeout 300 python - <<'PY'
import numpy as np
import numericalnim_b200 as nn
ctx = nn.default_context()
n = 1 << 24
y0 = 8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(n) / n)
g = nn.newVector(y0)
for name, rhs in (("builtin", nn.rhsLorenz96(8.0)), ("from source", nn.rhsJitStencil("((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, [], [8.0]))):
    for meth in ("tsit54", "vern65"):
        s = nn.Solver(meth, rhs, g, 1e12, nn.newODEoptions(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8))
        s.advance(5)
        ctx.set("profile", 1); ctx.profile_reset()
        s.advance(20)
        p = ctx.profile_read()["fused"]; ctx.set("profile", 0)
        print(name, meth, "us/launch", round(1e3 * p["ms"] / max(1, p["launches"]), 1), "GB/s", round(p["bytes"] / max(p["ms"], 1e-9) / 1e6, 1), "launches", p["launches"])
        s.close()
PY
