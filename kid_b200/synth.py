"""Seeded synthetic column domains for the Thompson step (SURVEY.md §8d "Synthetic inputs").

Everything is produced with torch tensor ops so the same code fills a CPU tensor (tests, the CPU
baseline sample) or a CUDA tensor (bench, inputs resident in HBM).  Randomness is a counter-based
integer hash of the GLOBAL column id, so a shard [c0, c0+n) of a domain equals the same columns
of the whole domain and results do not depend on how the domain is split over GPUs.

Layout: COL_FASTEST, every field is a float32 tensor (nz, ncol) - level-contiguous, columns
fastest - which is the device layout of the CUDA path.

Domain recipe (BASELINE.json configs 4/5): CONUS-like convective sounding (1000 hPa / 303 K /
16 g/kg at the surface, 6.5 K/km to a 210 K tropopause), nz = 60 levels of 250 m.  About 30 % of
the columns are cloudy, chosen by thresholding a smooth low-wavenumber 2-D field so that cloud
systems are spatially coherent as in a real convective scene; cloudy columns are one of
{shallow warm, deep convective, stratiform ice over rain, cirrus} with log-uniform peak contents.
"""
import math
import torch

SEED = 20261018
FIELDS = ("qv", "qc", "qi", "qr", "qs", "qg", "ni", "nr", "t")
_M64 = (1 << 64) - 1


def _i64(v):
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, s):
    return (x >> s) & ((1 << (64 - s)) - 1)


def hash_uniform(col_ids, stream, seed=SEED):
    """splitmix64 of (column id, stream) -> float32 uniform in [0, 1).  col_ids: int64 tensor."""
    x = col_ids * _i64(0x9E3779B97F4A7C15) + _i64((seed * 0x632BE59BD9B4E019 + stream * 0xD1B54A32D192ED03))
    x = (x ^ _lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _i64(0x94D049BB133111EB)
    x = x ^ _lsr(x, 31)
    return (_lsr(x, 40)).to(torch.float32) * (1.0 / (1 << 24))


def sounding(nz=60, dz=250.0, t_sfc=303.0, p_sfc=1.0e5, lapse=6.5e-3, t_min=210.0):
    """1-D base state at level mid-points: returns (z, T, p) float64 tensors of nz."""
    z = (torch.arange(nz, dtype=torch.float64) + 0.5) * dz
    T = torch.clamp(t_sfc - lapse * z, min=t_min)
    g, Rd = 9.81, 287.04
    z_trop = (t_sfc - t_min) / lapse
    p_trop = p_sfc * (t_min / t_sfc) ** (g / (Rd * lapse))
    p = torch.where(z <= z_trop, p_sfc * (T / t_sfc) ** (g / (Rd * lapse)),
                    p_trop * torch.exp(-g * (z - z_trop) / (Rd * t_min)))
    return z, T, p


def _qvs_liquid(p, T):
    # Bolton-type approximation, only used to place humidity; the scheme uses its own RSLF/RSIF.
    es = 611.2 * torch.exp(17.67 * (T - 273.15) / (T - 29.65))
    es = torch.minimum(es, 0.15 * p)
    return 0.622 * es / (p - es)


def _smooth_field(ix, iy, nx, ny, stream, nmodes=12, seed=SEED):
    """Sum of low-wavenumber cosine modes with hashed phases: coherent 2-D pattern in [-1, 1]-ish."""
    f = torch.zeros_like(ix, dtype=torch.float32)
    two_pi = 2.0 * math.pi
    for m in range(nmodes):
        h = hash_uniform(torch.tensor([m], dtype=torch.int64), 1000 + stream, seed)
        kx = 1 + int(float(hash_uniform(torch.tensor([m], dtype=torch.int64), 2000 + stream, seed)) * 6)
        ky = 1 + int(float(hash_uniform(torch.tensor([m], dtype=torch.int64), 3000 + stream, seed)) * 6)
        ph = float(h) * two_pi
        f += torch.cos(two_pi * (kx * ix.to(torch.float32) / nx + ky * iy.to(torch.float32) / ny) + ph)
    return f / math.sqrt(nmodes / 2.0)


def _bump(z, lo, hi):
    """Smooth vertical shape: 0 outside [lo, hi], sin^2 inside.  z: (nz,1), lo/hi: (1,ncol)."""
    s = torch.clamp((z - lo) / torch.clamp(hi - lo, min=1.0), 0.0, 1.0)
    return torch.sin(math.pi * s) ** 2


def make_domain(ncol, nz=60, col0=0, nx=None, device="cpu", seed=SEED, dz=250.0, cloudy_fraction=0.30,
                coherent=True, structure_block=1):
    """Columns [col0, col0+ncol) of a domain whose rows are nx columns wide (default 1024, or ncol
    when smaller).  Returns (state dict of (nz, ncol) float32 tensors, p (nz, ncol), dz (nz,))."""
    dev = torch.device(device)
    if nx is None:
        nx = 1024 if ncol + col0 >= 1024 else max(1, ncol + col0)
    ids = torch.arange(col0, col0 + ncol, dtype=torch.int64, device=dev)
    ix, iy = ids % nx, ids // nx
    ny = max(1024, 1)
    z1, T1, p1 = sounding(nz, dz)
    z = z1.to(dev, torch.float32).unsqueeze(1)
    # structure_block > 1 (experiments only): runs of that many neighbouring columns share their random numbers
    hid = ids if structure_block <= 1 else (ids // structure_block) * structure_block
    U = lambda s: hash_uniform(hid, s, seed).unsqueeze(0)          # (1, ncol)
    if coherent:
        sel = _smooth_field(ix, iy, nx, ny, 1, seed=seed)
        typ = _smooth_field(ix, iy, nx, ny, 2, seed=seed)
        # empirical quantile of a sum of 12 cosines (~N(0,1)): P(f > 0.524) ~ 0.30
        thr = {0.30: 0.524}.get(cloudy_fraction, math.sqrt(2.0) * _erfinv(1.0 - 2.0 * cloudy_fraction))
        cloudy = (sel > thr).unsqueeze(0)
        tsel = torch.clamp(0.5 + 0.35 * typ + 0.3 * (hash_uniform(ids, 7, seed) - 0.5), 0.0, 0.9999).unsqueeze(0)
    else:
        cloudy = (U(1) < cloudy_fraction)
        tsel = U(2)
    ctype = torch.floor(tsel * 4.0)   # 0 shallow warm, 1 deep convective, 2 stratiform, 3 cirrus

    T = T1.to(dev, torch.float32).unsqueeze(1) + 0.5 * _normal(ids, 11, seed).unsqueeze(0) * torch.ones_like(z)
    p = p1.to(dev, torch.float32).unsqueeze(1) * (1.0 + 0.002 * (U(12) - 0.5)) * torch.ones_like(z)
    qvs = _qvs_liquid(p, T)

    def logu(s, lo, hi):
        return torch.exp(math.log(lo) + U(s) * (math.log(hi) - math.log(lo)))

    zero = torch.zeros_like(T)
    is0, is1, is2, is3 = [(cloudy & (ctype == c)) for c in range(4)]
    # cloud water: shallow (0.8-3 km), deep (1-8 km), stratiform thin mixed layer (3-6 km)
    qc = torch.where(is0, logu(20, 1e-6, 2e-3) * _bump(z, 600.0 + 600.0 * U(21), 2000.0 + 1500.0 * U(22)), zero)
    qc = qc + torch.where(is1, logu(23, 1e-5, 2e-3) * _bump(z, 800.0 + 800.0 * U(24), 6500.0 + 2500.0 * U(25)), zero)
    qc = qc + torch.where(is2, logu(26, 1e-6, 3e-4) * _bump(z, 3000.0 + 1000.0 * U(27), 5500.0 + 1500.0 * U(28)), zero)
    # rain: from the surface up to 2-5 km
    rtop = 2000.0 + 3000.0 * U(30)
    rshape = torch.clamp((rtop - z) / 1500.0, 0.0, 1.0)
    qr = torch.where(is0, logu(31, 1e-7, 1e-3) * rshape, zero)
    qr = qr + torch.where(is1, logu(32, 1e-5, 5e-3) * torch.clamp((rtop + 1500.0 - z) / 2000.0, 0.0, 1.0), zero)
    qr = qr + torch.where(is2, logu(33, 1e-6, 1e-3) * rshape, zero)
    # ice aloft, snow through the mixed-phase layer, graupel in deep convection
    qi = torch.where(is1 | is2 | is3, logu(40, 1e-8, 5e-4) * _bump(z, 6000.0 + 2000.0 * U(41), 11000.0 + 3000.0 * U(42)), zero)
    qs = torch.where(is1 | is2, logu(43, 1e-7, 5e-3) * _bump(z, 3000.0 + 1500.0 * U(44), 9000.0 + 2500.0 * U(45)), zero)
    qg = torch.where(is1, logu(46, 1e-7, 5e-3) * _bump(z, 2500.0 + 1500.0 * U(47), 7500.0 + 2500.0 * U(48)), zero)
    qg = qg + torch.where(is2 & (U(49) < 0.3), logu(50, 1e-7, 5e-4) * _bump(z, 3000.0 + 1000.0 * U(51), 6000.0 + 1500.0 * U(52)), zero)
    floor = 2e-12
    qc, qr, qi, qs, qg = [torch.where(q > floor, q, zero) for q in (qc, qr, qi, qs, qg)]
    qc = torch.where(T > 236.0, qc, zero)   # no supercooled cloud below homogeneous freezing

    # numbers from the scheme's own size bounds, sizes growing with content (heavier rain has
    # bigger drops, denser ice cloud bigger crystals): rain mvd in [0.1, 2] mm, ice 20-250 um
    mvd = torch.clamp(1.0e-3 * torch.clamp(qr / 1.0e-3, min=1e-9) ** 0.25 * (0.6 + 0.8 * U(60)), 1.0e-4, 2.0e-3)
    lamr = 3.672 / mvd
    nr = torch.where(qr > 0, qr * lamr ** 3 / (math.pi * 1000.0), zero)
    xdi = torch.clamp(30.0e-6 * torch.clamp(qi / 1.0e-7, min=1e-9) ** 0.25 * (0.7 + 0.6 * U(61)), 20.0e-6, 250.0e-6)
    lami = 4.0 / xdi
    ni = torch.where(qi > 0, qi * lami ** 3 / (math.pi * 890.0), zero)

    # humidity: clear columns stay below ice saturation everywhere (they take the M:1540 exit);
    # cloudy columns are near water saturation inside liquid cloud, near ice saturation in ice.
    rh_clear = (0.75 - 0.45 * torch.clamp(z / 8000.0, 0.0, 1.0)) * (0.8 + 0.2 * U(70))
    rh_liq = 0.97 + 0.08 * U(71)
    esi_over_esl = torch.exp(torch.clamp(T - 273.15, max=0.0) * 0.0097)   # ~ qvsi/qvs
    rh_ice = esi_over_esl * (0.9 + 0.25 * U(72))
    rh_below = 0.55 + 0.4 * U(73)
    has_liq = qc > 0
    has_ice = (qi + qs + qg) > 0
    rh = torch.where(cloudy, rh_below * torch.ones_like(T), rh_clear * torch.ones_like(T))
    rh = torch.where(has_ice, torch.maximum(rh, rh_ice * torch.ones_like(T)), rh)
    rh = torch.where(has_liq, rh_liq * torch.ones_like(T), rh)
    qv = torch.clamp(rh * qvs, min=1e-7)

    state = {"qv": qv, "qc": qc, "qi": qi, "qr": qr, "qs": qs, "qg": qg, "ni": ni, "nr": nr, "t": T}
    state = {k: v.to(torch.float32).contiguous() for k, v in state.items()}
    dzv = torch.full((nz,), float(dz), dtype=torch.float32, device=dev)
    return state, p.to(torch.float32).contiguous(), dzv


def _normal(ids, stream, seed):
    u1 = torch.clamp(hash_uniform(ids, stream, seed), min=1e-7)
    u2 = hash_uniform(ids, stream + 500, seed)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)


def _erfinv(x):
    return float(torch.erfinv(torch.tensor(x, dtype=torch.float64)))


def deep_column(nz=60, dz=250.0):
    """BASELINE config 2: one mixed-phase deep column, every species present, deterministic.
    Returns numpy float32 arrays (dict by FIELDS), p, dz."""
    st, p, dzv = make_domain(4096, nz=nz, dz=dz, coherent=False, cloudy_fraction=1.0)
    # pick the deep-convective column with the largest total condensate
    tot = sum(st[k] for k in ("qc", "qr", "qi", "qs", "qg")).sum(0)
    has_all = (st["qc"].max(0).values > 1e-5) & (st["qr"].max(0).values > 1e-5) & (st["qi"].max(0).values > 1e-6) \
        & (st["qs"].max(0).values > 1e-5) & (st["qg"].max(0).values > 1e-5)
    tot = torch.where(has_all, tot, torch.zeros_like(tot))
    c = int(torch.argmax(tot))
    return ({k: v[:, c].numpy().copy() for k, v in st.items()}, p[:, c].numpy().copy(), dzv.numpy().copy())


def stats(state):
    """Presence fractions published next to every throughput number."""
    out = {}
    anyhyd = None
    for k in ("qc", "qr", "qi", "qs", "qg"):
        pres = state[k] > 1e-12
        out["cells_" + k] = float(pres.float().mean())
        anyhyd = pres if anyhyd is None else (anyhyd | pres)
    out["cells_any"] = float(anyhyd.float().mean())
    out["columns_any"] = float(anyhyd.any(0).float().mean())
    return out


def make_aerosols(state, p, seed=20261018):
    """Aerosol-aware inputs for the columns of `state` (dict of (nz, ncol) float32 arrays, torch or numpy) and pressure `p`:
    cloud droplet number nc [kg^-1] where there is cloud water, numbers of water-friendly and ice-friendly aerosols nwfa, nifa
    [kg^-1] decaying with height, vertical velocity w [m s^-1].  Seeded numpy; returns numpy float32 arrays (nc, nwfa, nifa, w)."""
    import numpy as np
    f = lambda a: a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)
    t, qv, qc, pp = f(state["t"]), f(state["qv"]), f(state["qc"]), f(p)
    nz, ncol = t.shape
    rng = np.random.default_rng(seed)
    rho = (np.float32(0.622) * pp / (np.float32(287.04) * t * (qv + np.float32(0.622)))).astype(np.float32)
    height = (1.0 - pp / pp.max()).astype(np.float32)                                   # 0 at the surface .. ~0.8 aloft
    nwfa = (10.0 ** rng.uniform(7.3, 9.3, (1, ncol)) * np.exp(-3.0 * height) / rho).astype(np.float32)
    nifa = (10.0 ** rng.uniform(3.5, 6.5, (1, ncol)) * np.exp(-2.0 * height) / rho).astype(np.float32)
    nc = np.where(qc > 1e-12, 10.0 ** rng.uniform(7.0, 8.9, (nz, ncol)) / rho, 0.0).astype(np.float32)
    w = (rng.normal(0.3, 1.5, (nz, ncol)) * (rng.random((1, ncol)) < 0.8)).astype(np.float32)
    return nc, nwfa, nifa, w
