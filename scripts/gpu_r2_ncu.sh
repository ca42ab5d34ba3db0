#!/bin/bash
# Round 2 ncu evidence (1 GPU): launch list of the default bench command, and --set full captures of the dominant kernels.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
B="python bench.py --steps 3 --warmup 3 --no-quad --no-jit --no-cpu-baseline --no-parity --e2e-reps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default_bench.csv $B > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fused_run_kernel" -s 3 -c 2 -o gpurun_out/r02_prof_fused_run $B --no-extra-configs > gpurun_out/ncu_fused_run.log 2>&1; tail -1 gpurun_out/ncu_fused_run.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"l96_attempt_kernel" -s 4 -c 1 -o gpurun_out/r02_prof_l96_attempt python bench.py --workload cfg3_tsit54_lorenz96_16M --steps 3 --warmup 3 --no-jit --no-quad --no-cpu-baseline --no-parity --no-extra-configs --e2e-reps 1 > gpurun_out/ncu_l96.log 2>&1; tail -1 gpurun_out/ncu_l96.log | cut -c1-200
timeout 400 ncu --set full --clock-control none -k regex:"fused_run_kernel" -s 1 -c 1 -o gpurun_out/r02_prof_fused_run_vern65 python bench.py --workload cfg4_vern65_diag_16M_per_gpu --steps 3 --warmup 3 --no-jit --no-quad --no-cpu-baseline --no-parity --no-extra-configs --e2e-reps 1 > gpurun_out/ncu_vern.log 2>&1; tail -1 gpurun_out/ncu_vern.log | cut -c1-200
timeout 400 ncu --set full --clock-control none -k regex:"stage_kernel|finish_kernel|ewise_kernel" -s 40 -c 13 -o gpurun_out/r02_prof_pipeline_attempt $B --no-extra-configs > gpurun_out/ncu_pipe.log 2>&1; tail -1 gpurun_out/ncu_pipe.log | cut -c1-200
ls -la gpurun_out | grep r02_ | cut -c1-120
