"""Kernel time of resident steps on the bench domain (CUDA events on the launching stream)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from kid_b200 import synth  # noqa: E402
from kid_b200.kidmp import Thompson, FIELDS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=1048576)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--col0", type=int, default=0)
ap.add_argument("--dt", type=float, default=10.0)
ap.add_argument("--cloudy", type=float, default=0.30)
ap.add_argument("--structure-block", type=int, default=1)
a = ap.parse_args()
th = Thompson(set_Nc=100.0, iiwarm=False, l_sediment=True)
st, p, dz = synth.make_domain(a.columns, nz=60, col0=a.col0, nx=1024, device="cuda", cloudy_fraction=a.cloudy, structure_block=a.structure_block)
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
d = th.diag()
print("ms per step:", ["%.2f" % x for x in ms], "active", d[6] / d[7], "Mcol/s %.1f" % (a.columns / min(ms) / 1e3))
th.close()
