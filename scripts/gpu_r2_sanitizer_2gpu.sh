set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
OUT=gpurun_out/sanitizer_report_2gpu.txt
: > $OUT
echo "== 2 GPUs without the tool" | tee -a $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 scripts/sanitizer_cases.py sharded 2>&1 | grep "sanitizer case\|Traceback\|Error" | tee -a $OUT
echo "== 2 GPUs: compute-sanitizer memcheck over both ranks (mailbox all-reduce, peer-read halo)" | tee -a $OUT
timeout 400 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 --print-limit 5 \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/sanitizer_cases.py sharded 2>&1 \
  | grep "sanitizer case\|ERROR SUMMARY\|Invalid\|========= .* at \|Traceback\|Error" | head -30 | cut -c1-250 | tee -a $OUT
echo "exit=${PIPESTATUS[0]}" | tee -a $OUT
echo "== 2 GPUs: racecheck" | tee -a $OUT
timeout 400 compute-sanitizer --tool racecheck --target-processes all --error-exitcode 7 --print-limit 5 \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 scripts/sanitizer_cases.py sharded 2>&1 \
  | grep "sanitizer case\|RACECHECK SUMMARY\|hazard\|Traceback\|Error" | head -30 | cut -c1-250 | tee -a $OUT
echo "exit=${PIPESTATUS[0]}" | tee -a $OUT
