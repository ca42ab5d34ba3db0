#!/bin/bash
# Fused kernel with compile-time sparsity patterns: parity, geometry matrix, bench, ncu; config-5 sweep on 1 GPU.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
echo "== tune"; timeout 900 python bench.py --tune 2>&1 | tee gpurun_out/tune.jsonl | grep fused_attempt | cut -c1-400
echo "== bench cfg2"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench.json | tail -1 | cut -c1-1800
echo "== bench cfg4 vern65 16M"; timeout 900 python bench.py --workload cfg4_vern65_diag_16M_per_gpu --no-cpu-baseline 2>&1 | tee gpurun_out/bench_cfg4_1gpu.json | tail -1 | cut -c1-600
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full: fused attempt kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_attempt -s 3 -c 2 -o gpurun_out/prof_fused_attempt \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
echo "== sweep (config 5, 1 GPU)"
timeout 1500 python bench.py --sweep --sweep-min 16 --sweep-max 28 --out gpurun_out/sweep_n1.json > gpurun_out/sweep_n1.log 2>&1; tail -4 gpurun_out/sweep_n1.log | cut -c1-300
ls gpurun_out
