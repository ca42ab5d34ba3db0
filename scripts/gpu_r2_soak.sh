#!/bin/bash
set -u
N=${1:-8}
export B200RK_JIT_CACHE=$PWD/.jitcache
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 scripts/soak_mailbox.py 2>&1 | grep "soak\|Error\|error\|Traceback" | head -10 | cut -c1-300
