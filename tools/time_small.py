"""Kernel time of small domains (KiD-size cases) for the launch-shape heuristic."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kid_b200 import synth
from kid_b200.kidmp import Thompson, FIELDS
th = Thompson(set_Nc=100.0, iiwarm=False, l_sediment=True)
print("graphs", os.environ.get("KIDMP_GRAPHS", "1"))
for ncol, nz in ((1, 60), (120, 120), (14400, 120), (65536, 60), (262144, 60)):
    st, p, dz = synth.make_domain(ncol, nz=nz, nx=1024, device="cuda", cloudy_fraction=1.0 if ncol < 65536 else 0.3, coherent=ncol >= 65536)
    ppt = torch.zeros((4, ncol), dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream(); torch.cuda.synchronize(); ms = []
    with torch.cuda.stream(s):
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            th.step_device(ncol, nz, 10.0, [st[k].data_ptr() for k in FIELDS], p.data_ptr(), dz.data_ptr(), ppt.data_ptr(), stream=s.cuda_stream)
            e1.record(); s.synchronize(); ms.append(e0.elapsed_time(e1))
    print("ncol %7d nz %3d  ms %.3f" % (ncol, nz, min(ms)))
th.close()
