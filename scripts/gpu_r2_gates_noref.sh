#!/bin/bash
# The round-end gates as the driver runs them (1 GPU): smoke, the full GPU suite, the default bench line, the reference arm.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -8 | cut -c1-300
timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 2> gpurun_out/gates_bench.err | grep '^{"metric"' | tee gpurun_out/gates_bench_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'parity', d['parity_check']['ok'], 'e2e', round(d['e2e']['value'], 1))
c3 = d.get('cfg3') or {}
print('cfg3', round(c3.get('value', 0), 1), 'from source', (c3.get('stencil_from_source') or {}))
print('cfg4', round((d.get('cfg4') or {}).get('value', 0), 1), 'jit', (d.get('jit_rhs') or {}).get('value'), 'quad', [(r['op'], round(r['GBps'])) for r in (d.get('trajectory_consumers') or {}).get('rows', [])])"
tail -2 gpurun_out/gates_bench.err | cut -c1-200
