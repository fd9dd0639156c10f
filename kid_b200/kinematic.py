"""Minimal kinematic host for BASELINE.json configs 1-3.

KiD's own driver (time loop, case definitions, advection, NetCDF output) is not part of the reference
mount; what IS fixed by the reference is the contract between the host and the scheme (I:54-97 and
I:198-245): the host hands the interface the state at the start of the step plus its own advective
and divergence tendencies, and gets microphysics tendencies back, which it adds.  This module is the
smallest host that honours that contract so that the same prescribed-flow case can be integrated
through the CUDA path and through the CPU oracle and their time series compared:

    state(n+1) = state(n) + dt * (d_adv + d_div + d_mphys)

Columns are independent: vertical first-order upwind advection by a prescribed updraft w(z, t) of
theta, qv and the seven hydrometeor moments; no horizontal transport (a 2-D case is nx independent
columns with different updraft strength / timing).  All host arithmetic is float32 numpy and is
identical for both back ends, which differ only in the `kid_interface` call.

Cases (shapes follow the KiD 1-D warm / deep and 2-D cumulus test cases in spirit, not in detail):
  warm1   : 1 column, nz = 60, dz = 50 m (0-3 km), constant 2 m/s updraft, iiwarm, dt = 1 s
  deep1   : 1 column, nz = 60, dz = 250 m (0-15 km), 3 m/s half-sine updraft for 40 min, mixed phase, dt = 5 s
  cu2d    : 120 columns x 120 levels (dz = 100 m), cosine-bell updraft across x, mixed phase, dt = 2 s
"""
import math

import numpy as np

HYD = ("qc", "qr", "nr", "qi", "ni", "qs", "qg")
P0 = np.float32(1.0e5)
R_ON_CP = np.float32(287.05 / 1005.0)
F32 = np.float32


def _qsat_liquid(p, T):
    es = 611.2 * np.exp(17.67 * (T - 273.15) / (T - 29.65))
    return 0.622 * es / (p - es)


class Case:
    def __init__(self, name, nx, nz, dz, dt, nsteps, iiwarm, set_Nc, t_sfc=297.0, rh_low=0.95, w_max=2.0, w_period=None,
                 w_top=None, lapse=6.0e-3):
        self.name, self.nx, self.nz, self.dt, self.nsteps = name, nx, nz, F32(dt), nsteps
        self.iiwarm, self.set_Nc = iiwarm, set_Nc
        self.dz = np.full(nz, dz, np.float32)
        z = (np.arange(nz) + 0.5) * dz
        self.z = z
        T = np.maximum(t_sfc - lapse * z, 205.0)
        g, Rd = 9.81, 287.05
        p = 1.0e5 * np.exp(-np.cumsum(g * dz / (Rd * T)) + 0.5 * g * dz / (Rd * T))
        exner = (p / 1.0e5) ** float(R_ON_CP)
        rh = np.where(z < 1500.0, rh_low, np.maximum(0.25, rh_low - (z - 1500.0) / 9000.0))
        qv = rh * _qsat_liquid(p, T)
        one = np.ones((nx, 1))
        self.exner = (one * exner).astype(np.float32)
        self.theta0 = (one * (T / exner)).astype(np.float32)
        self.qv0 = (one * qv).astype(np.float32)
        self.w_max, self.w_period, self.w_top = w_max, w_period, (w_top or z[-1])
        # per-column strength of the updraft (2-D case: cosine bell across x)
        x = (np.arange(nx) + 0.5) / nx
        self.w_x = (np.cos(np.pi * (x - 0.5)) ** 2 if nx > 1 else np.ones(1)).astype(np.float32)

    def w(self, t):
        """Updraft at the level interfaces below each level (nx, nz), m/s."""
        zi = np.arange(self.nz) * float(self.dz[0])
        shape = np.sin(np.pi * np.clip(zi / self.w_top, 0.0, 1.0)) if self.w_period else np.ones(self.nz)
        shape = np.where(zi <= self.w_top, shape, 0.0)
        amp = self.w_max * (math.sin(math.pi * min(t / self.w_period, 1.0)) if self.w_period else 1.0)
        return (self.w_x[:, None] * (amp * shape)[None, :]).astype(np.float32)


CASES = {
    "warm1": lambda: Case("warm1", 1, 60, 50.0, 1.0, 3600, True, 50.0, t_sfc=297.0, rh_low=0.97, w_max=2.0),
    "deep1": lambda: Case("deep1", 1, 60, 250.0, 5.0, 720, False, 100.0, t_sfc=300.0, rh_low=0.92, w_max=3.0,
                          w_period=2400.0, w_top=12000.0),
    "cu2d": lambda: Case("cu2d", 120, 120, 100.0, 2.0, 900, False, 100.0, t_sfc=298.0, rh_low=0.92, w_max=4.0,
                         w_period=1500.0, w_top=9000.0),
}


def _upwind(f, w, dz):
    """d f / dt by vertical upwind advection with w >= 0 at the interface below each level; the inflow at the
    bottom carries the bottom value (open lower boundary), f32 throughout."""
    below = np.concatenate([f[:, :1], f[:, :-1]], axis=1)
    return (-(w * (f - below)) / dz[None, :]).astype(np.float32)


def run(case, interface, nsteps=None, diag_every=1):
    """Integrate `case` with `interface(kid_dict, dt, p0, r_on_cp) -> dict` (the GPU binding's or the oracle's
    kid_interface).  Returns dict of time series: lwp, iwp (domain-mean kg/m^2), ppt (4, nt) domain-mean per step."""
    nx, nz, dt = case.nx, case.nz, case.dt
    st = {"theta": case.theta0.copy(), "qv": case.qv0.copy()}
    for m in HYD:
        st[m] = np.zeros((nx, nz), np.float32)
    zero = np.zeros((nx, nz), np.float32)
    out = {"lwp": [], "iwp": [], "ppt": [], "t": []}
    n = nsteps or case.nsteps
    p = (P0 * case.exner ** (F32(1.0) / R_ON_CP)).astype(np.float32)
    for it in range(n):
        t = float(it) * float(dt)
        w = case.w(t)
        kid = {"theta": st["theta"], "exner": case.exner, "qv": st["qv"], "dz": case.dz,
               "dtheta_adv": _upwind(st["theta"], w, case.dz), "dtheta_div": zero,
               "dqv_adv": _upwind(st["qv"], w, case.dz), "dqv_div": zero}
        for m in HYD:
            kid[m] = st[m]
            kid["d%s_adv" % m] = _upwind(st[m], w, case.dz)
            kid["d%s_div" % m] = zero
        ten = interface(kid, float(dt), float(P0), float(R_ON_CP))
        st["theta"] = (st["theta"] + dt * (kid["dtheta_adv"] + ten["dtheta_mphys"])).astype(np.float32)
        st["qv"] = (st["qv"] + dt * (kid["dqv_adv"] + ten["dqv_mphys"])).astype(np.float32)
        for m in (HYD[:3] if case.iiwarm else HYD):
            st[m] = (st[m] + dt * (kid["d%s_adv" % m] + ten["d%s_mphys" % m])).astype(np.float32)
        if it % diag_every == 0 or it == n - 1:
            T = st["theta"] * case.exner
            rho = F32(0.622) * p / (F32(287.04) * T * (st["qv"] + F32(0.622)))
            col = lambda q: float((q * rho * case.dz[None, :]).astype(np.float64).sum() / nx)
            out["lwp"].append(col(st["qc"] + st["qr"]))
            out["iwp"].append(col(st["qi"] + st["qs"] + st["qg"]))
            out["ppt"].append(ten["ppt"].astype(np.float64).sum(1) / nx)
            out["t"].append(t)
    out = {k: np.array(v) for k, v in out.items()}
    out["state"] = st
    return out
