#!/bin/bash
# Second GPU pass: parity suite after tolerance restatement, finish-grid matrix, bench, full ncu captures.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
echo "== tune"; timeout 600 python bench.py --tune 2>&1 | tee gpurun_out/tune.jsonl | tail -14
echo "== bench"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench.json | tail -3
echo "== bench vern65 16M"; timeout 900 python bench.py --workload cfg4_vern65_diag_16M_per_gpu --no-cpu-baseline 2>&1 | tee gpurun_out/bench_cfg4_1gpu.json | tail -3
echo "== bench tsit54 l96 16M"; timeout 900 python bench.py --workload cfg3_tsit54_lorenz96_16M --no-cpu-baseline 2>&1 | tee gpurun_out/bench_cfg3.json | tail -3
echo "== ncu full (stage m5 + finish)"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:stage_kernel<5|finish_kernel' -s 12 -c 4 -o gpurun_out/prof_stage5_finish \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
