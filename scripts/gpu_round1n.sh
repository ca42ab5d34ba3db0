#!/bin/bash
# 2-GPU box: sharded device loop (peer mailboxes inside the persistent kernel) + prefetching run kernel.
set -u
mkdir -p gpurun_out
echo "== pytest -x -q -m gpu"; timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "== multi gpu check (p2p)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -12
echo "== multi gpu check (nccl fallback)"; B200RK_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -3
summ='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]
print("value", round(d["value"],1), "us/step", round(1e3*d["ms_per_step"],2), "launches", d["gpu_launches"], "ach", round(r["achieved"]), "frac", round(r["frac"],3), "us/attempt", round(r.get("us_per_attempt",0),1), "e2e", round(d["e2e"]["value"],1), "pipeline", round(d["pipeline"]["value"],1) if d.get("pipeline") else None)'
echo "== bench cfg2 N=1"; timeout 900 python bench.py --gpus 1 --no-cpu-baseline 2>&1 | grep '^{"metric"' | tee gpurun_out/bench.json | python -c "$summ"
echo "== bench cfg4 N=1"; timeout 900 python bench.py --gpus 1 --no-cpu-baseline --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_cfg4_1gpu.json | python -c "$summ"
echo "== bench cfg4 N=1 host loop"; B200RK_DEVICE_LOOP=0 timeout 900 python bench.py --gpus 1 --no-cpu-baseline --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep '^{"metric"' | python -c "$summ"
echo "== bench cfg2 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_n2.json | python -c "$summ"
echo "== bench cfg2 N=2 host loop"; B200RK_DEVICE_LOOP=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 2>&1 | grep '^{"metric"' | python -c "$summ"
