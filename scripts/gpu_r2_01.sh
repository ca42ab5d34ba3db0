#!/bin/bash
# Round 2, GPU call 1 (1 GPU): first GPU measurements of the round-1 experimental kernels (l96 one-kernel attempt, finish prefetch, fused simpson)
#   1. first GPU run of what was written after round 1's GPU budget was spent (non-strict xfail tests + C++ demos)
#   2. cumsimpson: two-kernel path vs the single-pass kernel (knob fuse_simpson)
#   2c. config 3: the one-kernel Lorenz-96 attempt (knob fuse_stencil_attempt) vs the default path + its ncu capture
#   3. compute-sanitizer on TINY cases only (a 2-GPU memcheck pass over full-size tests ate round 1's last 13 minutes)
#   4. ncu --set full of the NVRTC-compiled device loop and of the cumtrapz scan
#   5. the round-end gates as the driver runs them
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
echo "== 1. experimental + C++ extras"; timeout 300 python -m pytest tests/test_zz_experimental_gpu.py tests/test_zz_cpp_extras_host.py -q -rxX -p no:cacheprovider 2>&1 | tail -12 | cut -c1-300
echo "== 2. cumsimpson two-kernel vs single-pass"
timeout 200 python bench.py --quad --out gpurun_out/quad_default.json 2>&1 | grep '"op": "cumsimpson"' | cut -c1-400
B200RK_FUSE_SIMPSON=1 timeout 200 python bench.py --quad --out gpurun_out/quad_fuse_simpson.json 2>&1 | grep '"op": "cumsimpson"' | cut -c1-400
echo "== 2b. finish kernel: default vs software-pipelined (pipeline leg of the bench: finish GB/s, steps/s)"
fin='import sys,json
d=json.loads(sys.stdin.read()); p=d["pipeline"]; f=p["roofline"]["finish_kernel"]
print("pipeline steps/s", round(p["value"],1), "finish GB/s", round(f["achieved"]), "us", round(f["avg_launch_us"],1))'
for cfg in "0 4 2" "1 4 2" "1 4 1" "1 2 2" "1 2 3"; do set -- $cfg
  echo "finish_prefetch=$1 vec_width=$2 finish_ctas_per_sm=$3"
  B200RK_FINISH_PREFETCH=$1 B200RK_VEC_WIDTH=$2 B200RK_FINISH_CTAS_PER_SM=$3 timeout 200 python bench.py --no-jit --no-quad --no-cpu-baseline --e2e-reps 1 2>&1 | grep '^{"metric"' | python -c "$fin"
done
echo "== 2c. config 3 (Tsit54, Lorenz-96, 2^24): default stage+stencil path vs the one-kernel attempt (knob fuse_stencil_attempt)"
l96='import sys,json
d=json.loads(sys.stdin.read()); a=d.get("l96_attempt") or {}
print("default steps/s", round(d["value"],1), "attempts/s", round(d["attempts_per_sec"],1), "| one-kernel attempt:", {k: (round(v,1) if isinstance(v,float) else v) for k,v in a.items() if k != "note"})'
timeout 400 python bench.py --workload cfg3_tsit54_lorenz96_16M --l96-attempt --no-jit --no-quad --no-cpu-baseline --e2e-reps 1 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_cfg3_l96_attempt.json | python -c "$l96" | cut -c1-700
B200RK_L96_ATTEMPT_PAIRS=1 timeout 400 python bench.py --workload cfg3_tsit54_lorenz96_16M --l96-attempt --no-jit --no-quad --no-cpu-baseline --e2e-reps 1 2>&1 | grep '^{"metric"' | python -c "$l96" | sed 's/^/512-wide tiles: /' | cut -c1-700
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"l96_attempt_kernel" -s 2 -c 1 -o gpurun_out/prof_l96_attempt \
  python bench.py --workload cfg3_tsit54_lorenz96_16M --l96-attempt --steps 3 --warmup 3 --no-jit --no-quad --no-cpu-baseline --e2e-reps 1 > gpurun_out/ncu_l96_attempt.log 2>&1; tail -1 gpurun_out/ncu_l96_attempt.log | cut -c1-200
ls gpurun_out
