"""Executed instructions per CUDA source line: joins the per-SASS-instruction counts of an ncu report (source page) with
the line table of the cubin (nvdisasm -g), instruction by instruction in program order.
    python tools/ncu_lines.py <report.ncu-rep> <library.so> <kernel substring, mangled> <kernel substring, ncu name> [top N]
Inlined code is attributed to the innermost line (helper functions show up under their own lines)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib, mangled, pretty = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
# line table of the kernel: list of (file, line) per instruction in order
lines, cur, on = [], ("?", 0), False
for l in dis:
    if l.startswith(".text."):
        on = mangled in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        lines.append((cur, l.split("*/", 1)[1].strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
counts, hdr, on = [], None, False
for row in csv.reader(src.split("\n")):
    if not row:
        continue
    if row[0] == "Kernel Name":
        if on and counts:
            break
        on = pretty in row[1]
        hdr = None
        continue
    if row[0] == "Address":
        hdr = {h: i for i, h in enumerate(row)}
        continue
    if on and hdr:
        counts.append((int(row[hdr["Instructions Executed"]] or 0), int(row[hdr["Thread Instructions Executed"]] or 0),
                       int(row[hdr["# Samples"]] or 0), row[hdr["Source"]].strip()))
print("instructions in cubin", len(lines), "in report", len(counts))
n = min(len(lines), len(counts))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for i in range(n):
    key = lines[i][0]
    for j in range(3):
        agg[key][j] += counts[i][j]
        tot[j] += counts[i][j]
print("total warp instr %.1f M, thread instr %.1f M, samples %d" % (tot[0] / 1e6, tot[1] / 1e6, tot[2]))
srcs = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kid_b200", "csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print("%5.1f%% inst %5.1f%% samp  %-18s %5d  %s" % (100 * v[0] / tot[0], 100 * v[2] / max(tot[2], 1), f, ln, text))
