#!/bin/bash
# The round-end gates exactly as the driver runs them (1 GPU): smoke, pytest -x -m gpu, bench (both arms); plus final ncu captures.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -x -q -m gpu"; timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "== bench reference arm"; ( time timeout 900 python bench.py --impl reference --gpus 1 --steps 30 --warmup 5 ) 2>&1 | tee gpurun_out/bench_reference.json | tail -5 | cut -c1-900
echo "== bench ours"; ( time timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 ) 2>&1 | tee gpurun_out/bench.json | tail -5 | cut -c1-600
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full: fused attempt kernel (prefetching, 4 CTAs/SM)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_attempt -s 20 -c 2 -o gpurun_out/prof_fused_attempt \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
ls gpurun_out
