"""Small driver for ncu: a few resident steps of the bench domain on one GPU.
    python tools/profile_step.py [--columns 262144] [--steps 3] [--col0 0]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from kid_b200 import synth  # noqa: E402
from kid_b200.kidmp import Thompson, FIELDS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=262144)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--col0", type=int, default=0)
ap.add_argument("--dt", type=float, default=10.0)
a = ap.parse_args()
th = Thompson(set_Nc=100.0, iiwarm=False, l_sediment=True)
st, p, dz = synth.make_domain(a.columns, nz=60, col0=a.col0, nx=1024, device="cuda")
ppt = torch.zeros((4, a.columns), dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
torch.cuda.synchronize()
for i in range(a.steps):
    th.step_device(a.columns, 60, a.dt, [st[k].data_ptr() for k in FIELDS], p.data_ptr(), dz.data_ptr(), ppt.data_ptr(),
                   stream=s.cuda_stream)
    s.synchronize()
    print("step", i, "active fraction", th.diag()[6] / a.columns)
th.close()
