#!/bin/bash
# 2 GPUs: sharded check incl. run-time compiled right-hand sides and the trajectory consumers; bench at N=2; sanitizer on the new kernels.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
echo "== multi gpu check (p2p)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -24 | cut -c1-300
echo "== multi gpu check (nccl fallback)"; B200RK_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -24 | cut -c1-300
echo "== bench cfg2 N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_n2_r1t.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'pipeline',round(d['pipeline']['value'],1))"
echo "== compute-sanitizer memcheck: jit + quadrature (small sizes)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_quadrature.py tests/test_gpu_rhs_from_source.py -q -p no:cacheprovider -x -k "not full_size and not device_loop_step_sequence and not single_step_matches" 2>&1 | tail -6 | cut -c1-300
echo "== compute-sanitizer racecheck: quadrature + fused jit attempt"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_quadrature.py tests/test_gpu_rhs_from_source.py -q -p no:cacheprovider -x -k "cumtrapz_bitwise or hermite_interpolate_bitwise or jit_fused_attempt" 2>&1 | tail -6 | cut -c1-300
