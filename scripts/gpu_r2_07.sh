#!/bin/bash
# Round 2, GPU call 7 (1 GPU): persistent l96_attempt_kernel with TMA bulk prefetch — parity, A/B of tile width x CTAs per SM, ncu.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
echo "== parity"
timeout 600 python -m pytest tests/test_gpu_fused_paths.py tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "l96 or lorenz or stencil" 2>&1 | tail -3 | cut -c1-300
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]
print("steps/s", round(d["value"],1), "attempts/s", round(d["attempts_per_sec"],1), "us/launch", round(r["avg_launch_us"],1), "frac", round(r["frac"],3))'
for cfg in "2 0" "1 0" "1 4" "2 2" "1 2" "1 3"; do set -- $cfg
  echo "pairs=$1 ctas_per_sm=$2"
  B200RK_L96_ATTEMPT_PAIRS=$1 B200RK_L96_CTAS_PER_SM=$2 timeout 300 python bench.py --workload cfg3_tsit54_lorenz96_16M --no-jit --no-quad --no-cpu-baseline --no-parity --no-extra-configs --e2e-reps 1 2>/dev/null | grep '^{"metric"' | python -c "$fmt"
done
echo "== vern65 on Lorenz-96 (2^24)"
timeout 300 python - <<'PY'
import numpy as np, time
import numericalnim_b200 as nn
ctx = nn.default_context()
n = 1 << 24
y0 = 8.0 + 0.01 * np.sin(2 * np.pi * 37 * np.arange(n) / n)
g = nn.newVector(y0)
for meth in ("dopri54", "tsit54", "vern65"):
    s = nn.Solver(meth, nn.rhsLorenz96(8.0), g, 1e12, nn.newODEoptions(absTol=1e-6, relTol=1e-6, dtMax=1.0, dtMin=1e-8))
    s.advance(5)
    ctx.set("profile", 1); ctx.profile_reset()
    s.advance(20)
    p = ctx.profile_read()["fused"]; ctx.set("profile", 0)
    print(meth, "us/launch", round(1e3 * p["ms"] / p["launches"], 1), "GB/s", round(p["bytes"] / p["ms"] / 1e6, 1))
    s.close()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"l96_attempt_kernel" -s 2 -c 1 -o gpurun_out/prof_l96_attempt_v2 \
  python bench.py --workload cfg3_tsit54_lorenz96_16M --steps 3 --warmup 3 --no-jit --no-quad --no-cpu-baseline --no-parity --no-extra-configs --e2e-reps 1 > gpurun_out/ncu_l96_attempt_v2.log 2>&1; tail -1 gpurun_out/ncu_l96_attempt_v2.log | cut -c1-200
