#!/bin/bash
# 2-GPU pass: full parity suite (incl. the sharded test), multi-GPU check, bench at N=1 and N=2 (torchrun).
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
echo "== multi gpu check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v Warning | tail -8
echo "== bench N=2 cfg2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | grep -v Warning | tee gpurun_out/bench_n2.json | tail -2 | cut -c1-1800
echo "== bench N=2 cfg4 vern65"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep -v Warning | tee gpurun_out/bench_cfg4_n2.json | tail -2 | cut -c1-1800
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tee gpurun_out/bench_reference.json | tail -2
ls gpurun_out
