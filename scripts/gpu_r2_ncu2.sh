#!/bin/bash
# ncu --set full of the final Lorenz-96 attempt kernel (128-thread CTAs) and of the NVRTC-compiled stencil-from-source attempt kernel.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"l96_attempt_kernel|ustencil_attempt_kernel" -s 4 -c 1 -o gpurun_out/r02_prof_l96_attempt_final python bench.py --workload cfg3_tsit54_lorenz96_16M --steps 3 --warmup 3 --no-jit --no-quad --no-cpu-baseline --no-parity --no-extra-configs --e2e-reps 1 > gpurun_out/ncu_l96f.log 2>&1; tail -1 gpurun_out/ncu_l96f.log | cut -c1-200
timeout 400 ncu --set full --clock-control none -k regex:"ustencil_attempt_kernel" -s 4 -c 1 -o gpurun_out/r02_prof_ustencil_attempt python - > gpurun_out/ncu_ust.log 2>&1 <<'PY'
import numpy as np
import numericalnim_b200 as nn
n = 1 << 24
y0 = 8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(n) / n)
s = nn.Solver("tsit54", nn.rhsJitStencil("((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, [], [8.0]), nn.newVector(y0), 1e12, nn.newODEoptions(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8))
s.advance(8)
s.close()
PY
tail -1 gpurun_out/ncu_ust.log | cut -c1-200
ls -la gpurun_out | grep "r02_prof_\(l96_attempt_final\|ustencil\)"
