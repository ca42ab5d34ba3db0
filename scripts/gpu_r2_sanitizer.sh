#!/bin/bash
# compute-sanitizer on TINY cases only (n <= 5003, each under its own timeout): memcheck on every kernel family, racecheck and
# synccheck on the kernels whose threads interact. 1 GPU; pass "2" to add the 2-GPU mailbox / peer-halo case.
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
OUT=gpurun_out/sanitizer_report.txt
: > $OUT
run() {  # tool, timeout, cases...
  local tool=$1 to=$2; shift 2
  echo "== compute-sanitizer --tool $tool : $*" | tee -a $OUT
  timeout $to compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 5 python scripts/sanitizer_cases.py "$@" 2>&1 | grep -v "^=========$" | grep "sanitizer case\|ERROR SUMMARY\|RACECHECK SUMMARY\|Error\|error\|hazard\|Invalid\|========= .* at \|Traceback" | head -30 | cut -c1-250 | tee -a $OUT
  echo "exit=${PIPESTATUS[0]}" | tee -a $OUT
}
python scripts/sanitizer_cases.py pipeline fused device_loop l96 quadrature jit 2>&1 | tail -7 | tee -a $OUT   # without the tool first: results + NVRTC cache warm
run memcheck 170 pipeline fused
run memcheck 170 device_loop l96
run memcheck 170 quadrature jit
run racecheck 170 device_loop
run racecheck 200 l96
run racecheck 170 fused quadrature
run synccheck 170 l96 device_loop
if [ "${1:-1}" = "2" ]; then
  echo "== 2 GPUs: memcheck over both ranks (mailbox all-reduce, peer-read halo)" | tee -a $OUT
  timeout 280 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 --print-limit 5 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/sanitizer_cases.py sharded 2>&1 \
    | grep "sanitizer case\|ERROR SUMMARY\|Invalid\|========= .* at \|Traceback" | head -20 | cut -c1-250 | tee -a $OUT
fi
