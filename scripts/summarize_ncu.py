#!/usr/bin/env python
"""Condense ncu artefacts brought back in gpurun_out/ into small tracked files under profiles/.
  summarize_ncu.py launches <launches.csv> <out.csv>     per-kernel totals / shares of a launch list
  summarize_ncu.py full <report.ncu-rep> <out.csv>        key metrics of every captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us", "grid", "block"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], round(a[1], 1), round(100 * a[1] / tot, 2), round(a[1] / a[0], 2), a[2], a[3]])


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx] + ["dram_bytes_total [Mbyte]", "dram_GBps"])
        for r in rows[2:]:
            vals = [r[i] for i in idx]
            rd, wr, us = (float(r[hdr.index(k)]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
            scale = {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}
            rd *= scale[units[hdr.index("dram__bytes_read.sum")]]
            wr *= scale[units[hdr.index("dram__bytes_write.sum")]]
            if units[hdr.index("gpu__time_duration.sum")] == "ms":
                us *= 1e3
            w.writerow(vals + [round(rd + wr, 3), round((rd + wr) / us * 1e3, 1)])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
