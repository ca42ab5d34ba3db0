#!/bin/bash
# Second GPU call of round 2 (gpurun --gpus 2, ~5 box-minutes = 10 GPU-minutes): the sharded one-kernel Lorenz-96 attempt
# (one ncclSend/ncclRecv halo exchange per step) in both all-reduce modes, then config 3 over 2 GPUs with and without it.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
echo "== multi gpu check (peer mailboxes)"; run 29511 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -28 | cut -c1-300
echo "== multi gpu check (nccl all-reduce)"; B200RK_P2P=0 run 29514 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -28 | cut -c1-300
echo "== config 3 over 2 GPUs: default (stage / RHS / finish pipeline, halo per evaluation) vs the one-kernel attempt"
run 29512 bench.py --gpus 2 --workload cfg3_tsit54_lorenz96_16M --l96-attempt --no-jit --no-quad --no-cpu-baseline --e2e-reps 1 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_cfg3_n2_l96_attempt.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); a = d.get('l96_attempt') or {}
print('default steps/s', round(d['value'], 1), '| one-kernel attempt:', {k: (round(v, 1) if isinstance(v, float) else v) for k, v in a.items() if k != 'note'})" | cut -c1-700
