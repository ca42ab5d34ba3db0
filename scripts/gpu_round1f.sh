#!/bin/bash
# 2-GPU box: parity suite, sharded checks (incl. Lorenz-96 ring halo), bench N=1 / N=2 with spin read-back + prefetching fused kernel.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
echo "== multi gpu check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v -i warning | tail -12
echo "== tune fused"; timeout 900 python bench.py --tune 2>&1 | tee gpurun_out/tune.jsonl | grep fused_attempt | grep '"vec_width": 4' | cut -c1-330
echo "== bench cfg2 N=1"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench.json | tail -1 | cut -c1-1500
echo "== bench cfg2 N=1 no spin"; B200RK_SPIN_READBACK=0 timeout 900 python bench.py --no-cpu-baseline 2>&1 | tee gpurun_out/bench_nospin.json | tail -1 | cut -c1-400
echo "== bench cfg2 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 2>&1 | grep -v -i warning | tee gpurun_out/bench_n2.json | tail -1 | cut -c1-700
echo "== bench cfg3 l96 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload cfg3_tsit54_lorenz96_16M 2>&1 | grep -v -i warning | tee gpurun_out/bench_cfg3_n2.json | tail -1 | cut -c1-700
echo "== bench cfg3 l96 N=1"; timeout 900 python bench.py --workload cfg3_tsit54_lorenz96_16M --no-cpu-baseline 2>&1 | tee gpurun_out/bench_cfg3.json | tail -1 | cut -c1-500
ls gpurun_out
