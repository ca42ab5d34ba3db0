#!/bin/bash
# 8-GPU box, trimmed: sharded device loop at 8 ranks (check + cfg2 + cfg4 = BASELINE config 4) and N=4.
set -u
mkdir -p gpurun_out
echo "== multi gpu check (8 ranks)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -9
summ='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]
print("value", round(d["value"],1), "us/step", round(1e3*d["ms_per_step"],2), "launches", d["gpu_launches"], "ach", round(r["achieved"]), "us/attempt", round(r.get("us_per_attempt",0),1), "e2e", round(d["e2e"]["value"],1), "pipeline", round(d["pipeline"]["value"],1) if d.get("pipeline") else None, d["config"]["error_norm_allreduce"])'
for n in 4 8; do
echo "== bench cfg2 N=$n"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n 2>&1 | grep '^{"metric"' | tee gpurun_out/scale_cfg2_n$n.json | python -c "$summ"
done
echo "== bench cfg4 vern65 N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep '^{"metric"' | tee gpurun_out/scale_cfg4_n8.json | python -c "$summ"
echo "== reference arm under torchrun N=8 (rank 0 only)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 2>&1 | grep '^{"impl"' | cut -c1-200
