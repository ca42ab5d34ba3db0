#!/bin/bash
# Round 2, GPU call 8 (1 GPU): l96_attempt_kernel with the next row's prefix evaluated in the shared-memory latency window.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
timeout 600 python -m pytest tests/test_gpu_fused_paths.py tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "l96 or lorenz or stencil" 2>&1 | tail -2 | cut -c1-300
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]
print("steps/s", round(d["value"],1), "attempts/s", round(d["attempts_per_sec"],1), "us/launch", round(r["avg_launch_us"],1), "frac", round(r["frac"],3))'
for cfg in "2 0" "1 0" "1 3"; do set -- $cfg
  echo "pairs=$1 ctas_per_sm=$2"
  B200RK_L96_ATTEMPT_PAIRS=$1 B200RK_L96_CTAS_PER_SM=$2 timeout 300 python bench.py --workload cfg3_tsit54_lorenz96_16M --no-jit --no-quad --no-cpu-baseline --no-parity --no-extra-configs --e2e-reps 1 2>/dev/null | grep '^{"metric"' | python -c "$fmt"
done
