"""Resource use and opcode histogram of every kernel of kid_b200/libkidmp.so (cuobjdump -res-usage, nvdisasm): the static
evidence next to the ncu counters.   python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "kid_b200", "libkidmp.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.split("\n"):
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
ops = collections.defaultdict(collections.Counter)
fn = None
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
    if m and fn:
        ops[fn][m.group(1)] += 1
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
print("# Static summary of the kernels of libkidmp.so (sm_100a SASS)\n")
print("`python tools/sass_summary.py`; registers / stack / static shared memory from `cuobjdump -res-usage`, instruction counts from `nvdisasm`")
print("(static counts: the out-of-line `dpow` / `dexp` / `dlog` bodies are part of the kernel they were linked into).\n")
print("| kernel | registers | stack B | smem B | SASS instructions | DFMA+DMUL+DADD | MUFU | F2F | LDG | STG | BAR | top opcodes |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for fn in sorted(ops, key=lambda f: -sum(ops[f].values())):
    if fn not in usage:
        continue
    c = ops[fn]
    tot = sum(c.values())
    g = lambda *names: sum(v for k, v in c.items() if k in names)
    top = ", ".join("%s %d" % kv for kv in c.most_common(5))
    name = demangle(fn).replace("kidmp::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
    name = name[:name.rfind("(")] if name.endswith(")") else name
    print("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %s |" % (name, *usage[fn], tot, g("DFMA", "DMUL", "DADD"), g("MUFU"), g("F2F"),
                                                                            g("LDG"), g("STG"), g("BAR"), top))
print("\nNo tensor-core (`UTCMMA` / `HMMA`) or TMA (`UTMALDG` / `UBLKCP`) instruction: nothing on this path is a dense contraction or a tile")
print("(DESIGN.md section 8); the 256-bit global accesses of sm_100 (`LDG.E.256` / `STG.E.256`) carry the hand-off records.")
