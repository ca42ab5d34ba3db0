#!/bin/bash
# BASELINE.json config 5: fused stage-combine bandwidth sweep N = 2^16 .. 2^28 (global length, sharded over the ranks) at N GPUs.
set -u
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --sweep --sweep-min 16 --sweep-max 28 --sweep-iters 50 --out gpurun_out/r02_sweep_config5_n1.json 2>&1 | tail -3 | cut -c1-300
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --sweep --sweep-min 16 --sweep-max 28 --sweep-iters 50 --out gpurun_out/r02_sweep_config5_n$N.json 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3 | cut -c1-300
fi
python - <<PY
import json
d = json.load(open("gpurun_out/r02_sweep_config5_n$N.json"))
for r in d["rows"]:
    if r["kernel"] in ("stage_m5", "finish_dopri54") and r["log2n"] in (16, 20, 23, 26, 28):
        print(r["log2n"], r["kernel"], round(r["GBps"]), "GB/s", round(r["frac_of_peak_per_gpu"], 3), "per-GPU frac", round(r["us_per_launch"], 1), "us")
    if r["kernel"].startswith("solver") and r["log2n"] in (16, 23, 26):
        print(r["log2n"], r["kernel"], round(r["steps_per_sec"]), "steps/s")
PY
