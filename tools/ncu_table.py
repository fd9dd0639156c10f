"""Pivot an `ncu --csv` launch list (one row per kernel launch and metric) into one line per launch.
    python tools/ncu_table.py gpurun_out/launches.csv [--step N launches per step]"""
import csv
import sys
from collections import OrderedDict

rows = OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    key = (int(r["ID"]), r["Kernel Name"])
    rows.setdefault(key, {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
names = []
for d in rows.values():
    for m in d:
        if m not in names:
            names.append(m)
short = {"gpu__time_duration.sum": "us", "smsp__inst_executed.sum": "Minst", "launch__registers_per_thread": "regs",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%", "smsp__issue_active.avg.pct": "issue%",
         "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes", "dram__bytes_read.sum": "rdMB",
         "dram__bytes_write.sum": "wrMB", "lts__t_sector_hit_rate.pct": "L2hit", "l1tex__t_sector_hit_rate.pct": "L1hit"}


def val(m, v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return float("nan")
    if m == "gpu__time_duration.sum":
        return x / 1e3 if u in ("ns", "nsecond") else (x * 1e3 if u.startswith("ms") else x)
    if m == "smsp__inst_executed.sum":
        return x / 1e6
    if m.startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return x * mult / 1e6
    return x


print("%4s %-44s" % ("id", "kernel") + "".join("%9s" % short.get(m, m[:8]) for m in names))
for (i, k), d in rows.items():
    kn = k.split("(")[0].replace("kidmp::", "")
    print("%4d %-44s" % (i, kn[:44]) + "".join("%9.1f" % val(m, *d[m]) if m in d else "%9s" % "-" for m in names))
