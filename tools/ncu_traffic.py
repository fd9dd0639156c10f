"""profiles/r02_traffic.json from an ncu launch list (CSV with dram__bytes_read.sum, dram__bytes_write.sum and
gpu__time_duration.sum per launch): the kernels of the LAST complete step (from its k_classify on), their DRAM bytes and
device times.  bench.py prints `bytes_per_step` as roofline.traffic.
    python tools/ncu_traffic.py profiles/r02_ncu_launches.csv profiles/r02_traffic.json"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
rows = OrderedDict()
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    rows.setdefault((int(r["ID"]), r["Kernel Name"]), {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])


def num(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3,
                "msecond": 1e3}.get(u, 1)


keys = list(rows)
starts = [i for i, (_, n) in enumerate(keys) if "k_classify" in n]
first = starts[-1]
step = keys[first:]
if not any("k_diag_reduce" in n for _, n in step):          # the capture ended inside the last step: take the one before
    first, last = starts[-2], starts[-1]
    step = keys[first:last]
per, total, us = OrderedDict(), 0.0, 0.0
for key in step:
    d = rows[key]
    name = key[1].split("(")[0].replace("void ", "").replace("kidmp::", "")
    b = num(*d["dram__bytes_read.sum"]) + num(*d["dram__bytes_write.sum"])
    t = num(*d["gpu__time_duration.sum"])
    e = per.setdefault(name, {"MB": 0.0, "us": 0.0})
    e["MB"] += b / 1e6
    e["us"] += t
    total += b
    us += t
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {"bytes_per_step": total, "sum_of_kernel_us": us, "source": src, "commit": commit,
       "workload": "1 048 576 columns x 60 levels, bench domain, one step (ncu: cold caches, kernels serialised)",
       "per_kernel_MB": {k: round(v["MB"], 1) for k, v in per.items()}, "per_kernel_us": {k: round(v["us"], 1) for k, v in per.items()}}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
