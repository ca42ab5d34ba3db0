#!/bin/bash
# 1 GPU: stencil-fused Lorenz-96 stage kernel: parity + cfg3 bench (fused vs pipeline) + ncu of the new kernel.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
echo "== bench cfg3 fused stencil"; timeout 900 python bench.py --workload cfg3_tsit54_lorenz96_16M --no-cpu-baseline 2>&1 | tee gpurun_out/bench_cfg3.json | tail -1 | cut -c1-2600
echo "== bench cfg3 pipeline"; B200RK_FUSE_STENCIL=0 timeout 900 python bench.py --workload cfg3_tsit54_lorenz96_16M --no-cpu-baseline 2>&1 | tee gpurun_out/bench_cfg3_nofuse.json | tail -1 | cut -c1-500
echo "== ncu full stage_l96"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_l96 -s 12 -c 6 -o gpurun_out/prof_stage_l96 \
  python bench.py --workload cfg3_tsit54_lorenz96_16M --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > gpurun_out/ncu_full3.log 2>&1; tail -1 gpurun_out/ncu_full3.log
ls gpurun_out
