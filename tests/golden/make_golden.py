"""Generate tests/golden/*.npz from the CPU oracle (run here, in the build container).

The reference ships no golden vectors and cannot be compiled (Fortran, no compiler): these
fixtures pin the ORACLE against silent drift and give the GPU tests inputs/outputs that do not
need the oracle library at run time.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kid_b200 import synth  # noqa: E402
from oracle.oracle import Oracle, FIELDS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # 48 columns: 32 all-cloudy mixed-phase + 16 from the 30 % cloudy coherent domain, nz = 60
    a, pa, dz = synth.make_domain(32, nz=60, cloudy_fraction=1.0, coherent=False)
    b, pb, _ = synth.make_domain(16, nz=60, col0=5000, nx=1024)
    st = {k: np.concatenate([a[k].numpy(), b[k].numpy()], axis=1) for k in FIELDS}
    p = np.concatenate([pa.numpy(), pb.numpy()], axis=1)
    out = {}
    for tag, kw, dt in (("mixed_dt10", dict(set_Nc=100.0, iiwarm=False), 10.0),
                        ("mixed_dt60", dict(set_Nc=100.0, iiwarm=False), 60.0),
                        ("warm_dt10", dict(set_Nc=50.0, iiwarm=True), 10.0),
                        ("nosed_dt10", dict(set_Nc=300.0, iiwarm=False, l_sediment=False), 10.0)):
        o = Oracle(**kw)
        s = {k: v.copy() for k, v in st.items()}
        ppt = o.step(dt, s, p.copy(), dz.numpy())
        for k in FIELDS:
            out["%s/%s" % (tag, k)] = s[k]
        out["%s/ppt" % tag] = ppt
        if tag == "mixed_dt10":
            for name in ("crg", "cgg", "csg", "cig", "cre", "cge", "cse", "cie", "scalars", "offsets"):
                out["const/" + name] = o.get(name)
            for name, idx in (("Dr", [0, 99]), ("Ds", [0, 99]), ("Dg", [0, 99]), ("Di", [0, 99]), ("t_Nc", [0, 99])):
                out["const/" + name] = o.get(name)[idx]
            # table checksums and a strided sample of every table
            for name in ("tcg_racg", "tmr_racg", "tcr_gacr", "tnr_racg", "tcs_racs1", "tmr_racs2", "tcr_sacr2", "tnr_sacr2",
                         "tpi_qrfz", "tpg_qrfz", "tni_qcfz", "tps_iaus", "tpi_ide", "t_Efrw", "t_Efsw"):
                t = o.get(name).ravel(order="F")
                out["table/%s/sum" % name] = np.array([t.sum(), np.abs(t).max(), float((t != 0).sum())])
                out["table/%s/sample" % name] = t[:: max(1, t.size // 257)][:257]
        o.close()
    for k in FIELDS:
        out["in/" + k] = st[k]
    out["in/p"] = p
    out["in/dz"] = dz.numpy()
    np.savez_compressed(os.path.join(HERE, "columns.npz"), **out)
    print("wrote", os.path.join(HERE, "columns.npz"), os.path.getsize(os.path.join(HERE, "columns.npz")), "bytes")


if __name__ == "__main__":
    main()
