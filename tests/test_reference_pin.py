"""The pin of the oracle to the reference itself: oracle/_ref/ref_driver is the UNMODIFIED Fortran
reference compiled with the stub host modules of oracle/ref/kid_stubs.f90.  It needs a Fortran compiler, which neither the
build container nor the GPU boxes of this pool have (profiles/r02_fortran_probe.txt), so these tests skip there; on a
machine with gfortran `make -C oracle/ref` builds the driver and they compare the C++ restatement with it column by column,
and write tests/golden/ref_columns.npz for the machines without a compiler."""
import os

import numpy as np
import pytest

from oracle.ref import ref
from oracle.oracle import FIELDS

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_columns.npz")


def test_stub_modules_cover_the_reference_imports():
    """Every name the two reference files import from the absent KiD host modules is defined by the stubs (checked on the
    text: the list below is M:19-23 and I:11-17).  Runs everywhere: it keeps the recipe honest without a compiler."""
    stubs = open(os.path.join(os.path.dirname(ref.__file__), "kid_stubs.f90")).read().lower()
    need = {"typekind": ["wp"], "switches": ["l_sediment", "l_reuse_thompson_lookup"], "diagnostics": ["save_dg", "i_dgtime"],
            "namelists": ["iiwarm", "set_nc"], "parameters": ["nx", "nz", "dt", "num_h_moments", "num_h_bins", "nspecies",
                                                              "h_names", "mom_units", "max_char_len"],
            "physconst": ["p0", "r_on_cp", "pi"],
            "column_variables": ["theta", "dtheta_adv", "dtheta_div", "dtheta_mphys", "exner", "qv", "dqv_adv", "dqv_div",
                                 "dqv_mphys", "dz", "hydrometeors", "dhydrometeors_adv", "dhydrometeors_div",
                                 "dhydrometeors_mphys"]}
    for mod, names in need.items():
        a = stubs.index("module " + mod)
        body = stubs[a:stubs.index("end module " + mod)]
        for n in names:
            assert n in body, (mod, n)


def test_table_checksum_matches_the_driver_formula():
    v = np.arange(1, 5001, dtype=np.float64) * 0.5
    s, w = ref.table_checksum(v)
    j = list(range(1, 5001, 997))
    assert s == v.sum() and w == sum(v[q - 1] * ((q % 1009) + 1) for q in j)


@pytest.mark.skipif(not ref.available() and ref.build() is None, reason="no Fortran compiler / reference sources here")
def test_oracle_equals_the_fortran_reference_on_the_golden_columns(oracle_mixed):
    from kid_b200 import synth
    st, p, dz = synth.make_domain(256, nz=60, cloudy_fraction=1.0, coherent=False)
    a = {k: np.ascontiguousarray(v.numpy().T) for k, v in st.items()}           # (nx, nz)
    pk = np.ascontiguousarray(p.numpy().T)
    got, ppt, tables = ref.run_columns(a, pk, dz.numpy(), 10.0, set_Nc=100.0)
    mine = {k: v.copy() for k, v in a.items()}
    ppt_o = oracle_mixed.step(10.0, mine, pk, dz.numpy(), layout="k_fastest")
    for k in FIELDS:
        den = np.maximum(np.abs(got[k]), 1e-12 if k.startswith("q") else 1e-3)
        assert (np.abs(mine[k] - got[k]) / den).max() <= 1e-5, k
    assert np.allclose(ppt_o, ppt, rtol=1e-5, atol=1e-12)
    for name in ("tcg_racg", "tmr_racs1", "tpi_qrfz", "tps_iaus", "t_Efrw"):
        s, w = ref.table_checksum(oracle_mixed.get(name).ravel(order="F"))
        assert abs(s - tables[name][0]) <= 1e-9 * abs(tables[name][0]) and abs(w - tables[name][1]) <= 1e-9 * abs(tables[name][1]), name
    np.savez_compressed(GOLD, p=pk, dz=dz.numpy(), ppt=ppt, **{"in_" + k: a[k] for k in FIELDS}, **{"out_" + k: got[k] for k in FIELDS})


@pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/ref_columns.npz is written on a machine with a Fortran compiler")
def test_oracle_equals_the_committed_reference_columns(oracle_mixed):
    g = np.load(GOLD)
    mine = {k: g["in_" + k].copy() for k in FIELDS}
    ppt_o = oracle_mixed.step(10.0, mine, g["p"], g["dz"], layout="k_fastest")
    for k in FIELDS:
        den = np.maximum(np.abs(g["out_" + k]), 1e-12 if k.startswith("q") else 1e-3)
        assert (np.abs(mine[k] - g["out_" + k]) / den).max() <= 1e-5, k
    assert np.allclose(ppt_o, g["ppt"], rtol=1e-5, atol=1e-12)


@pytest.mark.skipif(not ref.available() and ref.build() is None, reason="no Fortran compiler / reference sources here")
def test_oracle_equals_the_fortran_reference_with_the_aerosol_switch_on(oracle_mixed):
    """The aerosol-aware half: ref_driver mode 3 sets the module variable is_aerosol_aware (M:28) before thompson_init and passes
    nc1d, nwfa1d, nifa1d, w1d to mp_thompson; the oracle's restatement of the switch must give the same columns."""
    from kid_b200 import synth
    st, p, dz = synth.make_domain(256, nz=60, cloudy_fraction=1.0, coherent=False)
    nc, nwfa, nifa, w = synth.make_aerosols(st, p)
    a = {k: np.ascontiguousarray(v.numpy().T) for k, v in st.items()}           # (nx, nz)
    pk = np.ascontiguousarray(p.numpy().T)
    T = lambda x: np.ascontiguousarray(x.T)
    got, ppt, (gnc, gnwfa, gnifa), _ = ref.run_columns_aero(a, T(nc), T(nwfa), T(nifa), T(w), pk, dz.numpy(), 10.0, set_Nc=100.0)
    mine = {k: v.numpy().copy() for k, v in st.items()}                        # (nz, nx)
    ppt_o = oracle_mixed.step_aero(10.0, mine, nc, nwfa, nifa, p.numpy().copy(), w, dz.numpy())
    want = dict({k: got[k].T for k in FIELDS}, nc=gnc.T, nwfa=gnwfa.T, nifa=gnifa.T)
    have = dict(mine, nc=nc, nwfa=nwfa, nifa=nifa)
    for k in FIELDS + ("nc", "nwfa", "nifa"):
        den = np.maximum(np.abs(want[k]), 1e-12 if k.startswith("q") else 1e-3)
        assert (np.abs(have[k] - want[k]) / den).max() <= 1e-5, k
    assert np.allclose(ppt_o, ppt, rtol=1e-5, atol=1e-12)
