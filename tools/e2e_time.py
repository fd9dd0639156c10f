"""End-to-end time of kidmp_step with pinned host arrays (H2D + kernels + D2H every step), the path bench.py reports as e2e.
    KIDMP_PIPE_CHUNK=131072 python tools/e2e_time.py [--columns N] [--steps K]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from kid_b200 import synth  # noqa: E402
from kid_b200.kidmp import Thompson, FIELDS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=1048576)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
th = Thompson(set_Nc=100.0, iiwarm=False, l_sediment=True)
st, p, dz = synth.make_domain(a.columns, nz=60, nx=1024, device="cuda")
host = {k: torch.empty((60, a.columns), dtype=torch.float32).pin_memory() for k in FIELDS}
for k in FIELDS:
    host[k].copy_(st[k])
hp = torch.empty((60, a.columns), dtype=torch.float32).pin_memory()
hp.copy_(p)
del st, p
hs = {k: host[k].numpy() for k in FIELDS}
hpn, hdz = hp.numpy(), dz.cpu().numpy()
th.step(10.0, hs, hpn, hdz)
ms = []
for _ in range(a.steps):
    t0 = time.perf_counter()
    th.step(10.0, hs, hpn, hdz)
    ms.append((time.perf_counter() - t0) * 1e3)
gb = a.columns * 60 * 4 * 10 / 1e9
print("pipe_chunk", os.environ.get("KIDMP_PIPE_CHUNK", "default"), "zerocopy", os.environ.get("KIDMP_ZEROCOPY", "1"),
      "ms", " ".join("%.1f" % x for x in ms), "H2D GB/s %.1f" % (gb / (min(ms) * 1e-3)))
th.close()
