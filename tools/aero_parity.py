"""Aerosol-aware step (is_aerosol_aware = .true.): parity of kidmp_step_device_aero with the oracle on a few domains, and the
device time of the step on the bench domain.   python tools/aero_parity.py [--columns 1048576]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from kid_b200 import synth  # noqa: E402
from kid_b200.kidmp import Thompson, FIELDS  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
import parity_util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=1048576)
a = ap.parse_args()
g = Thompson(set_Nc=100.0, iiwarm=False, l_sediment=True)
o = Oracle(set_Nc=100.0, iiwarm=False, l_sediment=True)
names = FIELDS + ("nc", "nwfa", "nifa")
for ncol, dt, dz, frac in ((8192, 10.0, 250.0, 0.3), (8192, 60.0, 100.0, 0.6), (4096, 1.0, 250.0, 1.0)):
    st, p, dzv = synth.make_domain(ncol, nz=60, nx=1024, dz=dz, col0=777, cloudy_fraction=frac, coherent=False)
    nc, nwfa, nifa, w = synth.make_aerosols(st, p)
    ref = {k: v.numpy().copy() for k, v in st.items()}
    rn = [nc.copy(), nwfa.copy(), nifa.copy()]
    dev = {k: v.cuda() for k, v in st.items()}
    dn = [torch.from_numpy(x).cuda() for x in (nc, nwfa, nifa, w)]
    dp, ddz = p.cuda(), dzv.cuda()
    ppt = torch.zeros((4, ncol), dtype=torch.float32, device="cuda")
    o.step_aero(dt, ref, rn[0], rn[1], rn[2], p.numpy().copy(), w, dzv.numpy().copy())
    g.step_device_aero(ncol, 60, dt, [dev[k].data_ptr() for k in FIELDS], dn[0].data_ptr(), dn[1].data_ptr(), dn[2].data_ptr(),
                       dp.data_ptr(), dn[3].data_ptr(), ddz.data_ptr(), ppt.data_ptr())
    g.sync()
    got = {k: dev[k].cpu().numpy() for k in FIELDS}
    got.update(nc=dn[0].cpu().numpy(), nwfa=dn[1].cpu().numpy(), nifa=dn[2].cpu().numpy())
    want = dict(ref, nc=rn[0], nwfa=rn[1], nifa=rn[2])
    s = parity_util.compare_states(got, want, names)
    al = s["_all"]
    print("ncol %d dt %g dz %g cloudy %.1f: cells %d bit-identical %.5f within 1e-5 %.6f outside %d per field %s" % (
        ncol, dt, dz, frac, al["n"], al["exact_frac"], al["within_frac"], al["n"] - al["within"],
        {f: s[f]["n"] - s[f]["within"] for f in names if s[f]["n"] - s[f]["within"]}), flush=True)
ncol = a.columns
st, p, dzv = synth.make_domain(ncol, nz=60, nx=1024, device="cuda")
nc, nwfa, nifa, w = synth.make_aerosols(st, p)
dn = [torch.from_numpy(x).cuda() for x in (nc, nwfa, nifa, w)]
ppt = torch.zeros((4, ncol), dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
torch.cuda.synchronize()
ms = []
with torch.cuda.stream(s):
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.step_device_aero(ncol, 60, 10.0, [st[k].data_ptr() for k in FIELDS], dn[0].data_ptr(), dn[1].data_ptr(), dn[2].data_ptr(),
                           p.data_ptr(), dn[3].data_ptr(), dzv.data_ptr(), ppt.data_ptr(), stream=s.cuda_stream)
        e1.record()
        s.synchronize()
        ms.append(e0.elapsed_time(e1))
print("aerosol-aware step, %d columns x 60, bench domain: ms per step %s" % (ncol, " ".join("%.2f" % x for x in ms)))
g.set_option("timing", 1)
with torch.cuda.stream(s):
    g.step_device_aero(ncol, 60, 10.0, [st[k].data_ptr() for k in FIELDS], dn[0].data_ptr(), dn[1].data_ptr(), dn[2].data_ptr(),
                       p.data_ptr(), dn[3].data_ptr(), dzv.data_ptr(), ppt.data_ptr(), stream=s.cuda_stream)
    s.synchronize()
print("   kernels", " ".join("%s %.3f" % kv for kv in g.last_kernel_ms().items()))
print("   stats", g.step_stats())
g.close(); o.close()
