"""Run the CUDA path and the CPU oracle on the same seeded domains and print a parity report.

Usage (GPU box):  python tools/parity_report.py [--ncol 4096] [--out gpurun_out/parity.json]
The oracle is used here as the checker only.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from kid_b200 import synth  # noqa: E402
from kid_b200.kidmp import Thompson, FIELDS  # noqa: E402
from oracle.oracle import Oracle, TABLE_SHAPES  # noqa: E402
import parity_util  # noqa: E402


def tables_report(g, o):
    rep = {}
    for name in TABLE_SHAPES:
        a = g.get(name)
        b = o.get(name).ravel(order="F")
        denom = np.maximum(np.abs(b), 1e-300)
        rel = np.where(a == b, 0.0, np.abs(a - b) / denom)
        rep[name] = {"n": int(a.size), "exact": int((a == b).sum()), "max_rel": float(rel.max()),
                     "nonzero": int((b != 0).sum())}
    return rep


def consts_report(g, o):
    rep = {}
    for name in ("cre", "crg", "cse", "csg", "cge", "cgg", "cie", "cig", "ccg2", "ocg1", "scalars", "offsets", "Dr", "Ds",
                 "Dg", "Di", "Dc", "t_Nc", "dtr"):
        a, b = g.get(name), o.get(name).ravel()
        rep[name] = bool(np.array_equal(a, b))
    return rep


def domain_report(g, o, ncol, dt, col0=0, cloudy_fraction=0.3, coherent=True, nz=60):
    st, p, dz = synth.make_domain(ncol, nz=nz, col0=col0, cloudy_fraction=cloudy_fraction, coherent=coherent)
    ref = {k: v.numpy().copy() for k, v in st.items()}
    got = {k: v.numpy().copy() for k, v in st.items()}
    pn, dzn = p.numpy().copy(), dz.numpy().copy()
    t0 = time.time()
    ppt_ref = o.step(dt, ref, pn, dzn)
    t_cpu = time.time() - t0
    t0 = time.time()
    ppt_got = g.step(dt, got, pn, dzn)
    t_gpu = time.time() - t0
    got["ppt"], ref["ppt"] = ppt_got, ppt_ref
    parity_util.FLOOR["ppt"] = 1e-12
    stt = parity_util.compare_states(got, ref, FIELDS + ("ppt",))
    stt["_time"] = {"cpu_s": t_cpu, "gpu_e2e_s": t_gpu, "kernel_ms": g.last_step_ms()}
    stt["_stats"] = synth.stats(st)
    return stt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncol", type=int, default=4096)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rep = {}
    for tag, kw in (("mixed", dict(set_Nc=100.0, iiwarm=False)), ("warm", dict(set_Nc=50.0, iiwarm=True))):
        g = Thompson(**kw)
        o = Oracle(**kw)
        r = {"table_build_ms": g.table_build_ms, "consts": consts_report(g, o)}
        if tag == "mixed":
            r["tables"] = tables_report(g, o)
        for dt in (10.0, 60.0):
            r["domain_dt%g" % dt] = domain_report(g, o, a.ncol, dt, cloudy_fraction=1.0, coherent=False)
        r["domain_c30_dt10"] = domain_report(g, o, a.ncol, 10.0, cloudy_fraction=0.3, coherent=True)
        rep[tag] = r
        g.close(); o.close()
    txt = json.dumps(rep, indent=1)
    print(txt)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        open(a.out, "w").write(txt)


if __name__ == "__main__":
    main()
