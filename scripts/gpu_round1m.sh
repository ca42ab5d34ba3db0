#!/bin/bash
# Device loop as the default single-GPU path for element-local RHS: gates + ncu of fused_run_kernel.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -x -q -m gpu"; timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py --gpus 1 2>&1 | grep '^{"metric"' | tee gpurun_out/bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; p=d['pipeline']
print('value', round(d['value'],1), 'us/step', round(1e3*d['ms_per_step'],2), 'launches', d['gpu_launches'], 'path', d['path'][:40])
print('roofline', r['kernel'][:30], 'ach', round(r['achieved']), 'frac', round(r['frac'],3), 'us/attempt', round(r['us_per_attempt'],1), '| pipeline', round(p['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['ms_per_solve'])"
echo "== bench cfg4"; timeout 900 python bench.py --gpus 1 --no-cpu-baseline --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_cfg4_1gpu.json | cut -c1-200
echo "== ncu fused_run_kernel (10 steps)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_run -s 1 -c 1 -o gpurun_out/prof_fused_run \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/ncu_full4.log 2>&1; tail -1 gpurun_out/ncu_full4.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/bench_under_ncu.log 2>&1
ls gpurun_out
