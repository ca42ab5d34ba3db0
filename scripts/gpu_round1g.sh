#!/bin/bash
# 2-GPU box: in-kernel peer-mailbox all-reduce vs NCCL fallback; parity suite incl. new edge-case fixtures.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
echo "== multi gpu check (p2p mailboxes)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -12
echo "== multi gpu check (NCCL fallback)"; B200RK_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -12
echo "== bench cfg2 N=1"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tee gpurun_out/bench.json | tail -1 | cut -c1-420
echo "== bench cfg2 N=2 p2p"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 2>&1 | grep -v -i warning | tee gpurun_out/bench_n2.json | tail -1 | cut -c1-1300
echo "== bench cfg2 N=2 nccl"; B200RK_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 2>&1 | grep -v -i warning | tee gpurun_out/bench_n2_nccl.json | tail -1 | cut -c1-420
echo "== bench cfg4 N=2 p2p"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep -v -i warning | tee gpurun_out/bench_cfg4_n2.json | tail -1 | cut -c1-420
ls gpurun_out
