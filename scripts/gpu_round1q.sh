#!/bin/bash
# 1 GPU: run-time compiled right-hand sides (jit.cu) first, then the round-end gates, bench line, ncu launch list.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
echo "== jit tests"; timeout 900 python -m pytest tests/test_gpu_rhs_from_source.py -q -p no:cacheprovider 2>&1 | tail -40 | cut -c1-400
echo "== pytest -x -q -m gpu"; ( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -25 | cut -c1-400
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench ours"; ( time timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 ) 2>&1 | tee gpurun_out/bench_r1q.json | tail -5 | cut -c1-3000
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1q.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/bench_under_ncu_r1q.log 2>&1; tail -2 gpurun_out/bench_under_ncu_r1q.log | cut -c1-300
ls gpurun_out
