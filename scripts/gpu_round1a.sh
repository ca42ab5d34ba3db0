#!/bin/bash
# First GPU pass: smoke, parity tests, launch-geometry matrix, bench, ncu launch list + one full capture.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== tune"; timeout 600 python bench.py --tune 2>&1 | tee gpurun_out/tune.jsonl | tail -12
echo "== bench"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench.json | tail -3
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log
echo "== ncu full (stage kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 30 -c 4 -o gpurun_out/prof_stage \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
