#!/bin/bash
# A/B of the L2 hand-off on the fused attempt kernel + full parity with the auto policy.
set -u
mkdir -p gpurun_out
echo "== pytest gpu (auto L2 policy)"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for lg in 22 23; do for h in 0 1; do
  echo "== cfg2 log2n=$lg l2_hints=$h"
  B200RK_L2_HINTS=$h timeout 600 python bench.py --no-cpu-baseline --e2e-reps 1 --log2n $lg 2>&1 | grep '^{"metric"' | tee gpurun_out/l2f_lg${lg}_h${h}.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; p=d['pipeline']
print('fused steps/s', round(d['value'],1), 'us/step', round(1e3*d['ms_per_step'],2), 'kernel us', round(r['avg_launch_us'],1), 'GB/s', round(r['achieved']), '| pipeline steps/s', round(p['value'],1))"
done; done
echo "== cfg4 vern65 2^24 auto"; timeout 600 python bench.py --no-cpu-baseline --e2e-reps 1 --workload cfg4_vern65_diag_16M_per_gpu 2>&1 | grep '^{"metric"' | cut -c1-200
echo "== default bench"; timeout 600 python bench.py 2>&1 | grep '^{"metric"' | tee gpurun_out/bench.json | cut -c1-300
