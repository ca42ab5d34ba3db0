#!/bin/bash
# Round 2, GPU call 3 (1 GPU): the mailbox-based grid sum of the device loop (no grid barrier), new bench line.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
echo "== tests: device loop, fused paths, rhs from source (device loop cases)"
timeout 600 python -m pytest tests/test_gpu_device_loop.py tests/test_gpu_fused_paths.py -q -x -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_rhs_from_source.py -q -x -p no:cacheprovider -k "device_loop or sharded or solve" 2>&1 | tail -3 | cut -c1-300
echo "== bench N=1"
timeout 1200 python bench.py --steps 30 --warmup 5 2> gpurun_out/bench_r2_n1.err | grep '^{"metric"' | tee gpurun_out/bench_r2_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'], 1), 'repeats', [round(x, 3) for x in d['repeats']['ms_per_K_steps']], 'frac', round(d['roofline']['frac'], 3), 'us/attempt', round(d['roofline'].get('us_per_attempt', 0), 2))
print('parity', d['parity_check'])
print('pipeline', round(d['pipeline']['value'], 1), 'finish', d['pipeline']['roofline']['finish_kernel'])
for k in ('cfg3', 'cfg4'):
    c = d.get(k) or {}
    print(k, {x: (round(v, 2) if isinstance(v, float) else v) for x, v in c.items() if x not in ('roofline', 'config', 'path')}, 'frac', (c.get('roofline') or {}).get('frac'))
e = d['e2e']
print('e2e', round(e['value'], 1), 'ms/solve', round(e['ms_per_solve'], 3), 'pcie', e.get('pcie'), 'floor', e.get('pcie_floor_ms_per_solve'), 'resident', {k: e.get('rhs_resident', {}).get(k) for k in ('value', 'ms_per_solve', 'pcie_floor_ms_per_solve')}, 'numa', e.get('host_numa'))
print('jit', (d.get('jit_rhs') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'clocks', d['clocks'])
" | cut -c1-1200
tail -5 gpurun_out/bench_r2_n1.err | cut -c1-300
ls gpurun_out
