#!/bin/bash
# A/B: L2 eviction-priority hand-off (producer evict_last / streams evict_first) on the stage/RHS/finish pipeline.
set -u
mkdir -p gpurun_out
for lg in 22 23 24; do for h in 0 1; do
  echo "== pipeline log2n=$lg l2_hints=$h"
  B200RK_L2_HINTS=$h timeout 600 python bench.py --no-fuse --no-cpu-baseline --e2e-reps 1 --log2n $lg 2>&1 | grep '^{"metric"' | tee gpurun_out/l2_lg${lg}_h${h}.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('steps/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'stage GB/s', round(r['achieved']), 'us', round(r['avg_launch_us'],1), 'finish', round(r['finish_kernel']['achieved']), 'rhs GB/s', round(r['rhs_kernel']['achieved']))"
done; done
echo "== parity with hints on"; B200RK_L2_HINTS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "stage or golden or step or rk4 or fixed" 2>&1 | tail -3
echo "== ncu dram bytes per kernel, hints on"
B200RK_L2_HINTS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'stage_kernel|ewise_kernel|finish_kernel' -s 39 -c 13 --csv --log-file gpurun_out/l2_hints_on_attempt.csv \
  python bench.py --no-fuse --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > /dev/null 2>&1
B200RK_L2_HINTS=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'stage_kernel|ewise_kernel|finish_kernel' -s 39 -c 13 --csv --log-file gpurun_out/l2_hints_off_attempt.csv \
  python bench.py --no-fuse --steps 3 --warmup 3 --no-cpu-baseline --e2e-reps 1 > /dev/null 2>&1
python - <<'PY'
import csv
for tag in ("off","on"):
    rows=[r for r in csv.DictReader(l for l in open(f"gpurun_out/l2_hints_{tag}_attempt.csv") if not l.startswith("=="))]
    agg={}
    for r in rows:
        k=(r["ID"], r["Kernel Name"][:34]); agg.setdefault(k,{})[r["Metric Name"]]=(float(r["Metric Value"].replace(",","")), r["Metric Unit"])
    tot_r=tot_w=tot_t=0
    for k,v in agg.items():
        rd,ru=v["dram__bytes_read.sum"]; wr,wu=v["dram__bytes_write.sum"]; t,tu=v["gpu__time_duration.sum"]
        sc={"Mbyte":1,"Gbyte":1e3,"Kbyte":1e-3,"byte":1e-6}
        rd*=sc[ru]; wr*=sc[wu]; t=t/1e3 if tu=="ns" else t
        tot_r+=rd; tot_w+=wr; tot_t+=t
        print(tag, k[1], f"read {rd:7.1f} MB write {wr:7.1f} MB  {t:6.1f} us  hit {v['lts__t_sector_hit_rate.pct'][0]:.1f}%")
    print(tag, "TOTAL read", round(tot_r), "write", round(tot_w), "MB  time", round(tot_t,1), "us")
PY
