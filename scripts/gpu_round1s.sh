#!/bin/bash
# 1 GPU: bandwidth of the trajectory consumers + ncu launch list + one full capture of the cumtrapz kernel.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
echo "== quad bench 2^23"; timeout 600 python bench.py --quad --out gpurun_out/quad_2p23.json 2>&1 | tail -5 | cut -c1-700
echo "== quad bench 2^25, 17 points"; timeout 600 python bench.py --quad --log2n 25 --quad-points 17 --out gpurun_out/quad_2p25.json 2>&1 | tail -5 | cut -c1-700
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_quad.csv \
  python bench.py --quad --quad-iters 1 > gpurun_out/quad_under_ncu.log 2>&1; tail -1 gpurun_out/quad_under_ncu.log | cut -c1-200
echo "== ncu full: cumtrapz / simpson scan / hermite_many"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cumtrapz_kernel|simpson_scan_kernel|hermite_many_kernel" -s 6 -c 3 -o gpurun_out/prof_quad \
  python bench.py --quad --quad-iters 1 > gpurun_out/ncu_full_quad.log 2>&1; tail -1 gpurun_out/ncu_full_quad.log | cut -c1-200
ls -la gpurun_out | tail -8
