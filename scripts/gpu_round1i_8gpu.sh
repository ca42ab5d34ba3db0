#!/bin/bash
# 8-GPU box: sharded checks at 8 ranks, weak-scaling bench at N=1,2,4,8 (cfg2) and cfg4 (Vern65, 128M-dim over 8 GPUs), config-5 sweep at 8 GPUs.
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== multi gpu check (8 ranks)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -12
for n in 1 2 4 8; do
  echo "== bench cfg2 N=$n"
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline 2>&1 | grep '^{"metric"' | tee gpurun_out/scale_cfg2_n$n.json | cut -c1-140
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n 2>&1 | grep '^{"metric"' | tee gpurun_out/scale_cfg2_n$n.json | cut -c1-140; fi
done
for n in 1 8; do
  echo "== bench cfg4 vern65 N=$n"
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep '^{"metric"' | tee gpurun_out/scale_cfg4_n$n.json | cut -c1-140
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) bench.py --gpus $n --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep '^{"metric"' | tee gpurun_out/scale_cfg4_n$n.json | cut -c1-140; fi
done
echo "== bench cfg2 N=8 NCCL fallback"; B200RK_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 8 2>&1 | grep '^{"metric"' | tee gpurun_out/scale_cfg2_n8_nccl.json | cut -c1-140
echo "== bench cfg3 l96 N=8 (ring halo)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --workload cfg3_tsit54_lorenz96_16M 2>&1 | grep '^{"metric"' | tee gpurun_out/scale_cfg3_n8.json | cut -c1-140
echo "== sweep 8 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --sweep --sweep-min 18 --sweep-max 28 --sweep-step 2 --sweep-iters 50 --out gpurun_out/sweep_n8.json > gpurun_out/sweep_n8.log 2>&1; tail -3 gpurun_out/sweep_n8.log | cut -c1-200
ls gpurun_out
