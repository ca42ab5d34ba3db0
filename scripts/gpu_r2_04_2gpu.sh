#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): tagged-mailbox all-reduce (no fence, no grid barrier, no leader hop) — sharded parity + scaling at N = 2.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
N=${1:-2}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
echo "== multi gpu check (peer mailboxes)"; run 29511 scripts/multi_gpu_check.py 2>&1 | grep "multi-gpu\|FAIL\|Error\|error" | tail -30 | cut -c1-300
echo "== bench N=$N"
run 29512 bench.py --gpus $N --steps 30 --warmup 5 --no-quad --no-jit 2> gpurun_out/bench_r2_n$N.err | grep '^{"metric"' | tee gpurun_out/bench_r2_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'], 1), 'per gpu', round(d['value'] / d['n_gpus'], 1), 'us/step', round(1e3 * d['ms_per_step'], 2), 'repeats', [round(x, 3) for x in d['repeats']['ms_per_K_steps']], 'frac', round(d['roofline']['frac'], 3))
print('parity', {k: d['parity_check'][k] for k in ('ok', 'cases', 'max_rel', 'ranks')})
print('pipeline', round(d['pipeline']['value'], 1))
for k in ('cfg3', 'cfg4'):
    c = d.get(k) or {}
    print(k, {x: (round(v, 2) if isinstance(v, float) else v) for x, v in c.items() if x not in ('roofline', 'config', 'path', 'repeats')}, 'frac', (c.get('roofline') or {}).get('frac'))
e = d['e2e']
print('e2e', round(e['value'], 1), 'ms/solve', round(e['ms_per_solve'], 3), 'pcie', e.get('pcie'), 'resident', {k: e.get('rhs_resident', {}).get(k) for k in ('value', 'ms_per_solve')}, 'numa', e.get('host_numa'))
" | cut -c1-1200
tail -5 gpurun_out/bench_r2_n$N.err | cut -c1-300
echo "== cfg3 N=$N"
run 29513 bench.py --gpus $N --workload cfg3_tsit54_lorenz96_16M --steps 30 --warmup 5 --no-quad --no-jit --no-parity --no-extra-configs --e2e-reps 1 2>/dev/null | grep '^{"metric"' | tee gpurun_out/bench_r2_cfg3_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('cfg3 value', round(d['value'], 1), 'per gpu', round(d['value'] / d['n_gpus'], 1), 'frac', round(d['roofline']['frac'], 3), 'collectives', d['collectives'], 'attempts', d['attempts'], 'pipeline', round(d['pipeline']['value'], 1))"
ls gpurun_out | head -30
