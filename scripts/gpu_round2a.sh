#!/bin/bash
# First GPU call of the next round (1 GPU, ~6 GPU-minutes; every step under its own short timeout):
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
echo "== 3. compute-sanitizer, tiny cases"
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_quadrature.py -q -p no:cacheprovider -x -k "quadrature_errors or unsorted_input" 2>&1 | tail -4 | cut -c1-300
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x -k "error_norm_is_deterministic or test_lorenz96_rhs_bitwise" 2>&1 | tail -4 | cut -c1-300
echo "== 4. ncu full: NVRTC-compiled device loop, cumtrapz"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fused_run_kernel" -s 2 -c 1 -o gpurun_out/prof_jit_run \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-quad --e2e-reps 1 > gpurun_out/ncu_jit_run.log 2>&1; tail -1 gpurun_out/ncu_jit_run.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"cumtrapz_kernel" -s 3 -c 1 -o gpurun_out/prof_cumtrapz \
  python bench.py --quad --quad-iters 1 > gpurun_out/ncu_cumtrapz.log 2>&1; tail -1 gpurun_out/ncu_cumtrapz.log | cut -c1-200
echo "== 5. gates"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_r2a.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'jit', d.get('jit_rhs', {}).get('value'), 'quad', d.get('trajectory_consumers'))" | cut -c1-900
ls gpurun_out
