#!/bin/bash
# The driver's scaling run for one N, as the driver launches it (plus, at N = 8, the config-5 sweep).
set -u
mkdir -p gpurun_out
export B200RK_JIT_CACHE=$PWD/.jitcache
N=${1:-1}
summ='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; e=d["e2e"]
print("N", d["n_gpus"], "value", round(d["value"],1), "per gpu", round(d["value"]/d["n_gpus"],1), "us/step", round(1e3*d["ms_per_step"],2), "frac", round(r["frac"],3), "repeats", [round(x,3) for x in d["repeats"]["ms_per_K_steps"]])
print("  parity", d["parity_check"]["ok"], d["parity_check"]["cases"], d["parity_check"]["max_rel"], "pipeline", round(d["pipeline"]["value"],1))
for k in ("cfg3","cfg4"):
    c=d.get(k) or {}
    if c: print("  ", k, round(c.get("value",0),1), "frac", round((c.get("roofline") or {}).get("frac",0),3), "collectives", c.get("collectives"), "attempts", c.get("attempts"), c.get("error"), "from source", (c.get("stencil_from_source") or {}).get("value"))
print("  e2e", round(e["value"],1), "ms/solve", round(e["ms_per_solve"],3), "resident", round(e.get("rhs_resident",{}).get("value",0),1), "alt tstart", e.get("tstart_slot",{}).get("alternative_value"), "pcie", e.get("pcie"))
print("  clocks", d["clocks"])'
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 30 --warmup 5 2> gpurun_out/scale_r2_n1.err | grep '^{"metric"' | tee gpurun_out/scale_r2_n1.json | python -c "$summ"
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 30 --warmup 5 2> gpurun_out/scale_r2_n$N.err | grep '^{"metric"' | tee gpurun_out/scale_r2_n$N.json | python -c "$summ"
  echo "== cfg3 sharded N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N bench.py --gpus $N --workload cfg3_tsit54_lorenz96_16M --steps 30 --warmup 5 --no-quad --no-jit --no-parity --no-extra-configs --e2e-reps 1 2>/dev/null | grep '^{"metric"' | tee gpurun_out/scale_r2_cfg3_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('cfg3 N', d['n_gpus'], 'value', round(d['value'], 1), 'per gpu', round(d['value'] / d['n_gpus'], 1), 'frac', round(d['roofline']['frac'], 3), 'collectives', d['collectives'], 'attempts', d['attempts'])"
fi
tail -2 gpurun_out/scale_r2_n$N.err | cut -c1-200
if [ "$N" = "8" ] && [ "${2:-}" = "sweep" ]; then bash scripts/gpu_r2_sweep.sh 8; fi
