#!/bin/bash
# 1 GPU: trajectory consumers (quadrature.cu) parity, then the whole GPU suite.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
echo "== quadrature tests"; ( time timeout 900 python -m pytest tests/test_gpu_quadrature.py -q -p no:cacheprovider ) 2>&1 | tail -60 | cut -c1-600
echo "== pytest -x -q -m gpu"; ( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -12 | cut -c1-400
