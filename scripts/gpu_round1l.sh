#!/bin/bash
# Device-resident driver loop: parity, then the small-N end of the config-5 sweep with and without it.
set -u
mkdir -p gpurun_out
echo "== pytest device loop"; timeout 900 python -m pytest tests/test_gpu_device_loop.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15
echo "== pytest all gpu"; timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -5
echo "== sweep small N"; timeout 900 python bench.py --sweep --sweep-min 16 --sweep-max 23 --sweep-iters 30 --sweep-cpu-max 0 --sweep-solver-steps 200 --out gpurun_out/sweep_small.json 2>&1 | grep solver_ | cut -c1-200
echo "== bench default"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric"' | cut -c1-200
