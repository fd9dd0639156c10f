"""ORACLE - TEST INFRASTRUCTURE ONLY.
Python side of oracle/ref: builds (when a Fortran compiler and /root/reference exist) and runs oracle/_ref/ref_driver,
the UNMODIFIED reference scheme compiled with the stub host modules of kid_stubs.f90.

    probe()      -> {"compiler": path or None, "tried": [...]}     (recorded in bench.py's cpu_baseline)
    build()      -> path of the driver, or None when it cannot be built here
    run_columns(state, p, dz, dt, ...)    one mp_thompson call per column (U1 inputs), returns (state after, ppt[4, nx], tables)
    run_columns_aero(state, nc, nwfa, nifa, w, p, dz, dt, ...)   the same with is_aerosol_aware = .true. (M:28)
    run_interface(kid, dt, ...)           one mphys_thompson09_interfacen call, returns the KiD tendencies
Arrays are KiD's (k, i) order: numpy shape (nx, nz), float32.
"""
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
DRIVER = os.path.join(OUT, "ref_driver")
REFERENCE = os.environ.get("KID_REFERENCE", "/root/reference")
COMPILERS = ("gfortran", "flang", "flang-new", "nvfortran", "ifx", "ifort", "f95", "pgfortran")
FIELDS = ("qv", "qc", "qi", "qr", "qs", "qg", "ni", "nr", "t")
HYD = ("qc", "qr", "nr", "qi", "ni", "qs", "qg")
TABLES = ("t_Efrw", "tcg_racg", "tmr_racg", "tcr_gacr", "tmg_gacr", "tnr_racg", "tnr_gacr", "tcs_racs1", "tmr_racs1", "tcs_racs2",
          "tmr_racs2", "tcr_sacr1", "tms_sacr1", "tcr_sacr2", "tms_sacr2", "tnr_racs1", "tnr_racs2", "tnr_sacr1", "tnr_sacr2",
          "tpi_qcfz", "tni_qcfz", "tpi_qrfz", "tpg_qrfz", "tni_qrfz", "tnr_qrfz", "tps_iaus", "tni_iaus", "tpi_ide", "t_Efsw")


def probe():
    found = None
    for c in COMPILERS:
        p = shutil.which(c)
        if p:
            found = p
            break
    return {"compiler": found, "tried": list(COMPILERS)}


def available():
    return os.path.exists(DRIVER)


def build(force=False):
    """make -C oracle/ref; returns the driver path or None (no compiler, or no reference sources on this machine)."""
    if available() and not force:
        return DRIVER
    fc = probe()["compiler"]
    if not fc or not os.path.exists(os.path.join(REFERENCE, "module_mp_thompson09n.f90")):
        return None
    subprocess.check_call(["make", "-C", HERE, "FC=" + fc, "REF=" + REFERENCE])
    return DRIVER if available() else None


def table_checksum(values):
    """The two words ref_driver.f90::chk writes for a table given in Fortran element order."""
    v = np.asarray(values, np.float64).ravel()
    j = np.arange(1, v.size + 1, 997)
    return float(v.sum()), float((v[j - 1] * ((j % 1009) + 1)).sum())


def _run(header, arrays, nout):
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        with open(fin, "wb") as f:
            f.write(header)
            for a in arrays:
                f.write(np.ascontiguousarray(a, np.float32).tobytes())
        os.makedirs(os.path.join(OUT, "run_data"), exist_ok=True)
        subprocess.check_call([DRIVER, fin, fout], cwd=OUT)      # the reference opens run_data/rac[gs]_thompson09.data (M:3710)
        raw = np.fromfile(fout, np.uint8)
    body = raw[:nout * 4].view(np.float32)
    tw = raw[nout * 4:nout * 4 + 64 * 8].view(np.float64)
    tables = {n: (float(tw[2 * i]), float(tw[2 * i + 1])) for i, n in enumerate(TABLES)}
    tables["save_dg_calls"] = int(tw[63])
    return body, tables


def _header(mode, nx, nz, iiwarm, l_sediment, dt, set_Nc):
    return np.array([mode, nx, nz, int(iiwarm), int(l_sediment)], np.int32).tobytes() + np.array([dt, set_Nc], np.float32).tobytes()


def run_columns(state, p, dz, dt, set_Nc=100.0, iiwarm=False, l_sediment=True):
    nx, nz = np.asarray(state["t"]).shape
    arrays = [np.stack([np.asarray(state[k], np.float32) for k in FIELDS]), p, dz]      # (9, nx, nz) = Fortran (nz, nx, 9)
    body, tables = _run(_header(1, nx, nz, iiwarm, l_sediment, dt, set_Nc), arrays, 9 * nx * nz + 4 * nx)
    out = body[:9 * nx * nz].reshape(9, nx, nz)
    ppt = body[9 * nx * nz:].reshape(nx, 4).T.copy()                                     # Fortran (4, nx)
    return {k: out[i].copy() for i, k in enumerate(FIELDS)}, ppt, tables


def run_columns_aero(state, nc, nwfa, nifa, w, p, dz, dt, set_Nc=100.0, iiwarm=False, l_sediment=True):
    """mp_thompson per column with is_aerosol_aware = .true.; returns (state after, ppt[4, nx], (nc, nwfa, nifa) after, tables)."""
    nx, nz = np.asarray(state["t"]).shape
    arrays = [np.stack([np.asarray(state[k], np.float32) for k in FIELDS]), p,
              np.stack([np.asarray(a, np.float32) for a in (nc, nwfa, nifa, w)]), dz]
    body, tables = _run(_header(3, nx, nz, iiwarm, l_sediment, dt, set_Nc), arrays, 9 * nx * nz + 4 * nx + 3 * nx * nz)
    out = body[:9 * nx * nz].reshape(9, nx, nz)
    ppt = body[9 * nx * nz:9 * nx * nz + 4 * nx].reshape(nx, 4).T.copy()
    ae = body[9 * nx * nz + 4 * nx:].reshape(3, nx, nz)
    return {k: out[i].copy() for i, k in enumerate(FIELDS)}, ppt, tuple(ae[i].copy() for i in range(3)), tables


def run_interface(kid, dt, set_Nc=100.0, iiwarm=False, l_sediment=True):
    nx, nz = np.asarray(kid["theta"]).shape
    arrays = [kid[n] for n in ("theta", "dtheta_adv", "dtheta_div", "exner", "qv", "dqv_adv", "dqv_div", "dz")]
    for m in HYD:
        arrays += [kid[m], kid["d%s_adv" % m], kid["d%s_div" % m]]
    body, tables = _run(_header(2, nx, nz, iiwarm, l_sediment, dt, set_Nc), arrays, 9 * nx * nz)
    out = body.reshape(9, nx, nz)
    res = {"dtheta_mphys": out[0].copy(), "dqv_mphys": out[1].copy()}
    for i, m in enumerate(HYD):
        res["d%s_mphys" % m] = out[2 + i].copy()
    return res, tables
