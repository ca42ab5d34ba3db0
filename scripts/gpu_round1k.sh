#!/bin/bash
# 2-GPU box: verification of the split build (5 translation units): all parity tests, sharded checks, gates.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -x -q -m gpu"; timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "== multi gpu check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -12
echo "== bench"; timeout 900 python bench.py --gpus 1 2>&1 | grep '^{"metric"' | tee gpurun_out/bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; p=d['pipeline']
print('fused steps/s', round(d['value'],1), 'kernel us', round(r['avg_launch_us'],1), 'frac', round(r['frac'],3), '| pipeline steps/s', round(p['value'],1), 'stage frac', round(p['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'])"
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_n2.json | cut -c1-160
