"""Which cells differ from the oracle by more than 1e-5 (every outlier was in qc)?
Runs the six domains of the parity report on the GPU and the oracle, and prints every cell outside tolerance with its
inputs, both results and the saturation state it started from.   python tools/qc_outliers.py [--ncol 4096]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from kid_b200 import synth  # noqa: E402
from kid_b200.kidmp import Thompson, FIELDS  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
import parity_util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ncol", type=int, default=4096)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_qc_outliers.json"))
a = ap.parse_args()
g = Thompson(set_Nc=100.0, iiwarm=False, l_sediment=True)
o = Oracle(set_Nc=100.0, iiwarm=False, l_sediment=True)
cases = [dict(col0=0, cloudy_fraction=0.3, coherent=True, dt=10.0), dict(col0=300000, cloudy_fraction=0.3, coherent=True, dt=10.0),
         dict(col0=0, cloudy_fraction=1.0, coherent=False, dt=10.0), dict(col0=7, cloudy_fraction=1.0, coherent=False, dt=1.0),
         dict(col0=0, cloudy_fraction=0.6, coherent=False, dt=60.0), dict(col0=1000, cloudy_fraction=1.0, coherent=False, dt=20.0)]
rows = []
for ci, c in enumerate(cases):
    dt = c.pop("dt")
    st, p, dz = synth.make_domain(a.ncol, nz=60, **c)
    ref = {k: v.numpy().copy() for k, v in st.items()}
    got = {k: v.numpy().copy() for k, v in st.items()}
    inp = {k: v.numpy().copy() for k, v in st.items()}
    pn = p.numpy().copy()
    o.step(dt, ref, pn, dz.numpy())
    g.step(dt, got, pn, dz.numpy())
    nbad = {}
    for f in FIELDS:
        den = np.maximum(np.abs(ref[f].astype(np.float64)), parity_util.FLOOR[f])
        rel = np.abs(got[f].astype(np.float64) - ref[f]) / den
        rel[got[f] == ref[f]] = 0
        bad = np.argwhere(rel > 1e-5)
        nbad[f] = len(bad)
        for k, j in bad[:40]:
            t, pp, qv = float(inp["t"][k, j]), float(pn[k, j]), float(inp["qv"][k, j])
            qvs = orc.rslf(pp, t)
            rows.append({"case": ci, "dt": dt, "field": f, "k": int(k), "col": int(j), "got": float(got[f][k, j]), "ref": float(ref[f][k, j]),
                         "rel": float(rel[k, j]), "T": t, "p": pp, "qv": qv, "ssatw_in": qv / qvs - 1.0,
                         "in": {q: float(inp[q][k, j]) for q in ("qc", "qi", "qr", "qs", "qg")},
                         "got_all": {q: float(got[q][k, j]) for q in FIELDS}, "ref_all": {q: float(ref[q][k, j]) for q in FIELDS}})
    print("case", ci, "dt", dt, "cells outside 1e-5 per field:", nbad, flush=True)
json.dump(rows, open(a.out, "w"), indent=1)
for r in rows[:60]:
    print("case %d %s k=%2d col=%5d got %.6e ref %.6e rel %.2e  T %.2f ssatw_in %+.3e in %s" % (
        r["case"], r["field"], r["k"], r["col"], r["got"], r["ref"], r["rel"], r["T"], r["ssatw_in"],
        {q: "%.2e" % v for q, v in r["in"].items() if v}))
    print("        got t %.6f qv %.6e qc %.6e | ref t %.6f qv %.6e qc %.6e" % (
        r["got_all"]["t"], r["got_all"]["qv"], r["got_all"]["qc"], r["ref_all"]["t"], r["ref_all"]["qv"], r["ref_all"]["qc"]))
g.close(); o.close()
