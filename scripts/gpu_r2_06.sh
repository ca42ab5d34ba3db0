#!/bin/bash
# Round 2, GPU call 6 (1 GPU): full GPU suite after the reciprocal / sign-folding / mailbox changes, bench line, host topology.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
echo "== host topology"; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" | head -8; ls /sys/devices/system/node/ 2>/dev/null | head; nvidia-smi topo -m 2>/dev/null | head -12
echo "== full GPU suite"
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -8 | cut -c1-300
echo "== bench N=1"
timeout 1200 python bench.py --steps 30 --warmup 5 --no-quad 2> gpurun_out/bench_r2b_n1.err | grep '^{"metric"' | tee gpurun_out/bench_r2b_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'], 1), 'repeats', [round(x, 3) for x in d['repeats']['ms_per_K_steps']], 'frac', round(d['roofline']['frac'], 3), 'us/attempt', round(d['roofline'].get('us_per_attempt', 0), 2))
print('parity', d['parity_check']['ok'], d['parity_check']['max_rel'])
print('pipeline', round(d['pipeline']['value'], 1), 'finish', d['pipeline']['roofline']['finish_kernel'])
for k in ('cfg3', 'cfg4'):
    c = d.get(k) or {}
    print(k, {x: (round(v, 2) if isinstance(v, float) else v) for x, v in c.items() if x not in ('roofline', 'config', 'path')}, 'frac', (c.get('roofline') or {}).get('frac'))
print('jit', (d.get('jit_rhs') or {}).get('value'))
" | cut -c1-1200
tail -3 gpurun_out/bench_r2b_n1.err | cut -c1-300
