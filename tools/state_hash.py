"""SHA-256 of every state field after a few resident steps on the bench domain, plus the kernel time.

Used to show that an optimisation of the kernels is bit-exact: run it with KIDMP_LIB pointing at the
previous build of the library and again with the new one; the hashes must be identical.
"""
import argparse
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from kid_b200 import synth  # noqa: E402
from kid_b200.kidmp import Thompson, FIELDS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=1048576)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--dt", type=float, default=10.0)
ap.add_argument("--dz", type=float, default=250.0)
ap.add_argument("--warm", action="store_true")
ap.add_argument("--timing", type=int, default=0, help="after the timed steps, three more in timing mode 1 (kernels serialised) or 2 (normal schedule): ms of every kernel group")
a = ap.parse_args()
th = Thompson(set_Nc=100.0, iiwarm=a.warm, l_sediment=True)
st, p, dz = synth.make_domain(a.columns, nz=60, nx=1024, device="cuda", dz=a.dz)
ppt = torch.zeros((4, a.columns), dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
torch.cuda.synchronize()
ms = []
with torch.cuda.stream(s):
    for i in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        th.step_device(a.columns, 60, a.dt, [st[k].data_ptr() for k in FIELDS], p.data_ptr(), dz.data_ptr(), ppt.data_ptr(),
                       stream=s.cuda_stream)
        e1.record()
        s.synchronize()
        ms.append(e0.elapsed_time(e1))
h = hashlib.sha256()
for k in FIELDS:
    h.update(st[k].cpu().numpy().tobytes())
h.update(ppt.cpu().numpy().tobytes())
print("lib", os.environ.get("KIDMP_LIB", "in-tree"), "columns", a.columns, "steps", a.steps, "dt", a.dt, "dz", a.dz,
      "warm", a.warm, "sha256", h.hexdigest()[:32], "ms", " ".join("%.2f" % x for x in ms))
if a.timing:
    th.set_option("timing", a.timing)
    acc = {}
    with torch.cuda.stream(s):
        for i in range(3):
            th.step_device(a.columns, 60, a.dt, [st[k].data_ptr() for k in FIELDS], p.data_ptr(), dz.data_ptr(), ppt.data_ptr(),
                           stream=s.cuda_stream)
            s.synchronize()
            for k, v in th.last_kernel_ms().items():
                acc[k] = acc.get(k, 0.0) + v / 3
    print("   kernels", " ".join("%s %.3f" % kv for kv in acc.items()), "sum %.3f" % sum(acc.values()))
try:
    print("   stats", th.step_stats())
except Exception as e:      # a library of an earlier round
    pass
th.close()
