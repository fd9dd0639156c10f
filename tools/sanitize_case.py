"""Small mixed-phase + warm + ragged cases for compute-sanitizer (one tool per run)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kid_b200 import synth
from kid_b200.kidmp import Thompson, FIELDS
for kw, ncol, nz, dt in ((dict(set_Nc=100.0, iiwarm=False), 1000, 60, 60.0), (dict(set_Nc=50.0, iiwarm=True), 333, 37, 5.0),
                         (dict(set_Nc=100.0, iiwarm=False), 77, 120, 10.0)):
    th = Thompson(**kw)
    st, p, dz = synth.make_domain(ncol, nz=nz, cloudy_fraction=0.6, coherent=False)
    s = {k: v.numpy().copy() for k, v in st.items()}
    ppt = th.step(dt, s, p.numpy(), dz.numpy())
    kt = {k: np.ascontiguousarray(v.numpy().T) for k, v in st.items()}
    th.step(dt, kt, np.ascontiguousarray(p.numpy().T), dz.numpy(), layout="k_fastest")
    print(kw, ncol, nz, "ok", float(ppt.sum()), th.diag())
    th.close()
