// ORACLE — TEST INFRASTRUCTURE ONLY (see thompson_oracle.h).  PARITY UNPINNED by the reference.
//
// Line-by-line CPU restatement of /root/reference/module_mp_thompson09n.f90 ("M:") with Fortran
// expression semantics kept: default REAL = f32, DOUBLE PRECISION = f64, an expression is
// evaluated in the widest kind among the operands seen so far (left to right), NINT = round half
// away from zero, INT = truncate, real**integer = libgcc powi (repeated multiplication).
// Build: g++ -O2 -fno-fast-math -ffp-contract=off (no FMA contraction, like gfortran x86-64).
//
// Decisions on undefined behaviour of the reference: U1..U10, marked inline.
#include "thompson_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <unistd.h>
#include <string>
#include <vector>
#include <map>
#include <chrono>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef float  r4;
typedef double r8;

// ---- Fortran intrinsics -------------------------------------------------------------------
inline int nint_f(r4 x) { return (int)lroundf(x); }
inline int nint_d(r8 x) { return (int)lround(x); }
// real ** integer as libgcc __powisf2 / __powidf2 (what gfortran emits for 10.**nn, Dc**nu_c)
inline r4 powi_f(r4 x, int m) {
  unsigned n = m < 0 ? (unsigned)(-m) : (unsigned)m;
  r4 y = (n % 2) ? x : 1.0f;
  while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
  return m < 0 ? 1.0f / y : y;
}
inline r8 powi_d(r8 x, int m) {
  unsigned n = m < 0 ? (unsigned)(-m) : (unsigned)m;
  r8 y = (n % 2) ? x : 1.0;
  while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
  return m < 0 ? 1.0 / y : y;
}
inline r4 maxf(r4 a, r4 b) { return a > b ? a : b; }
inline r4 minf(r4 a, r4 b) { return a < b ? a : b; }
inline r8 maxd(r8 a, r8 b) { return a > b ? a : b; }
inline r8 mind(r8 a, r8 b) { return a < b ? a : b; }
inline r4 signf(r4 a, r4 b) { return std::signbit(b) ? -fabsf(a) : fabsf(a); }

// ---- PARAMETERs, M:30-204 -----------------------------------------------------------------
const r4 T_0 = 273.15f;
const r4 PI = 3.1415926536f;
const r4 rho_w = 1000.0f, rho_s = 100.0f, rho_g = 500.0f, rho_i = 890.0f;
const r4 Nt_c_max = 1999.E6f;
const r4 naIN1 = 0.5E6f;
const r4 mu_r = 0.0f, mu_g = 0.0f, mu_i = 0.0f;
const r4 mu_s = 0.6357f, Kap0 = 490.6f, Kap1 = 17.46f, Lam0 = 20.78f, Lam1 = 3.29f;
const r4 gonv_min = 1.E4f, gonv_max = 3.E6f;
const r4 am_r = PI * rho_w / 6.0f, bm_r = 3.0f;
const r4 am_s = 0.069f, bm_s = 2.0f;
const r4 am_g = PI * rho_g / 6.0f, bm_g = 3.0f;
const r4 am_i = PI * rho_i / 6.0f, bm_i = 3.0f;
const r4 av_r = 4854.0f, bv_r = 1.0f, fv_r = 195.0f;
const r4 av_s = 40.0f, bv_s = 0.55f, fv_s = 100.0f;
const r4 av_g = 442.0f, bv_g = 0.89f;
const r4 av_i = 1847.5f, bv_i = 1.0f;
const r4 bv_c = 2.0f;
const r4 C_cube = 0.5f, C_sqrd = 0.15f;
const r4 Ef_si = 0.05f, Ef_rs = 0.95f, Ef_rg = 0.75f, Ef_ri = 0.95f;
const r4 R1 = 1.E-12f, R2 = 1.E-6f, eps = 1.E-15f;
const r4 TNO = 5.0f, ATO = 0.304f;
const r4 rho_not = 101325.0f / (287.05f * 298.0f);
const r4 Sc = 0.632f;
const r4 HGFR = 235.16f;
const r4 Rv = 461.5f, oRv = 1.f / Rv, R = 287.04f, Cp = 1004.0f;
const r4 lsub = 2.834E6f, lvap0 = 2.5E6f, lfus = lsub - lvap0, olfus = 1.f / lfus;
const r4 xm0i = 1.E-12f, D0c = 1.E-6f, D0r = 50.E-6f, D0s = 200.E-6f, D0g = 250.E-6f;
const int IFDRY = 0;

enum { nbins = 100, nbc = 100, nbi = 100, nbr = 100, nbs = 100, nbg = 100 };
enum { ntb_c = 37, ntb_i = 64, ntb_r = 37, ntb_s = 28, ntb_g = 28, ntb_g1 = 28, ntb_r1 = 37,
       ntb_i1 = 55, ntb_t = 9, ntb_IN = 55 };

// table axes M:215-303 : digits 1..9 times decades, closed by one extra decade value
void fill_axis(r4* a, int n, const char* const* lits) { for (int i = 0; i < n; ++i) a[i + 1] = strtof(lits[i], 0); }
// literal text keeps the exact decimal->f32 conversion gfortran performs on "1.e-6" etc.
std::vector<std::string> axis_literals(int dec_lo, int dec_hi_incl_single) {
  std::vector<std::string> v;
  for (int d = dec_lo; d < dec_hi_incl_single; ++d)
    for (int m = 1; m <= 9; ++m) { char b[32]; snprintf(b, 32, "%d.e%d", m, d); v.push_back(b); }
  char b[32]; snprintf(b, 32, "1.e%d", dec_hi_incl_single); v.push_back(b);
  return v;
}
void make_axis(r4* a, int n, int dec_lo, int dec_hi) {
  std::vector<std::string> v = axis_literals(dec_lo, dec_hi);
  if ((int)v.size() != n) { fprintf(stderr, "oracle: axis size mismatch\n"); abort(); }
  for (int i = 0; i < n; ++i) a[i + 1] = strtof(v[i].c_str(), 0);
}

const r4 sa[11] = {0, 5.065339f, -0.062659f, -3.032362f, 0.029469f, -0.000285f,
                   0.31255f, 0.000204f, 0.003199f, 0.0f, -0.015952f};
const r4 sb[11] = {0, 0.476221f, -0.015896f, 0.165977f, 0.007468f, -0.000141f,
                   0.060366f, 0.000079f, 0.000594f, 0.0f, -0.003577f};
const r4 Tc[10] = {0, -0.01f, -5.f, -10.f, -15.f, -20.f, -25.f, -30.f, -35.f, -40.f};

// ---- gamma family, M:4530-4651 ---------------------------------------------------------------
r4 GAMMLN(r4 XX) {
  const r8 STP = 2.5066282746310005;
  const r8 COF[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                     -1.231739572450155, .1208650973866179E-2, -.5395239384953E-5};
  r8 X = XX, Y = X, TMP = X + 5.5;
  TMP = (X + 0.5) * log(TMP) - TMP;
  r8 SER = 1.000000000190015;
  for (int J = 0; J < 6; ++J) { Y = Y + 1.0; SER = SER + COF[J] / Y; }
  return (r4)(TMP + log(STP * SER / X));
}
r4 WGAMMA(r4 y) { return expf(GAMMLN(y)); }
void GCF(r4& GAMMCF, r4 A, r4 X, r4& GLN) {
  const int ITMAX = 100; const r4 gEPS = 3.E-7f, FPMIN = 1.E-30f;
  GLN = GAMMLN(A);
  r4 B = X + 1.f - A, C = 1.f / FPMIN, D = 1.f / B, H = D, AN, DEL;
  for (int I = 1; I <= ITMAX; ++I) {
    AN = -((r4)I * ((r4)I - A));
    B = B + 2.f;
    D = AN * D + B;
    if (fabsf(D) < FPMIN) D = FPMIN;
    C = B + AN / C;
    if (fabsf(C) < FPMIN) C = FPMIN;
    D = 1.f / D;
    DEL = D * C;
    H = H * DEL;
    if (fabsf(DEL - 1.f) < gEPS) break;
  }
  GAMMCF = expf(-X + A * logf(X) - GLN) * H;
}
void GSER(r4& GAMSER, r4 A, r4 X, r4& GLN) {
  const int ITMAX = 100; const r4 gEPS = 3.E-7f;
  GLN = GAMMLN(A);
  if (X <= 0.f) { GAMSER = 0.f; return; }
  r4 AP = A, SUM = 1.f / A, DEL = SUM;
  for (int N = 1; N <= ITMAX; ++N) {
    AP = AP + 1.f;
    DEL = DEL * X / AP;
    SUM = SUM + DEL;
    if (fabsf(DEL) < fabsf(SUM) * gEPS) break;
  }
  GAMSER = SUM * expf(-X + A * logf(X) - GLN);
}
r4 GAMMP(r4 A, r4 X) {
  r4 g = 0.f, GLN;
  if (X < 0.f || A <= 0.f) return 0.f;
  if (X < A + 1.f) { GSER(g, A, X, GLN); return g; }
  GCF(g, A, X, GLN); return 1.f - g;
}

// ---- saturation mixing ratios, M:4656-4717 ------------------------------------------------------
r4 RSLF(r4 P, r4 T) {
  const r4 C0 = .611583699E03f, C1 = .444606896E02f, C2 = .143177157E01f, C3 = .264224321E-1f,
           C4 = .299291081E-3f, C5 = .203154182E-5f, C6 = .702620698E-8f, C7 = .379534310E-11f,
           C8 = -.321582393E-13f;
  r4 X = maxf(-80.f, T - 273.16f);
  r4 ESL = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
  ESL = minf(ESL, P * 0.15f);
  return .622f * ESL / (P - ESL);
}
r4 RSIF(r4 P, r4 T) {
  const r4 C0 = .609868993E03f, C1 = .499320233E02f, C2 = .184672631E01f, C3 = .402737184E-1f,
           C4 = .565392987E-3f, C5 = .521693933E-5f, C6 = .307839583E-7f, C7 = .105785160E-9f,
           C8 = .161444444E-12f;
  r4 X = maxf(-80.f, T - 273.16f);
  r4 ESI = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
  ESI = minf(ESI, P * 0.15f);
  return .622f * ESI / (P - ESI);
}

// ---- the decade-mantissa table index, M:1762-1774 (f32 argument) and M:1824-1833 (f64) --------
int decade_idx_f(r4 x, int n2, int ntb) {
  int n0 = nint_f(log10f(x)), n = n0;
  for (int nn = n0 - 1; nn <= n0 + 1; ++nn) {
    n = nn;
    if ((x / powi_f(10.f, nn)) >= 1.0f && (x / powi_f(10.f, nn)) < 10.0f) break;
  }
  int idx = (int)(x / powi_f(10.f, n)) + 10 * (n - n2) - (n - n2);
  return std::max(1, std::min(idx, ntb));
}
int decade_idx_d(r8 x, int n2, int ntb) {
  int n0 = nint_d(log10(x)), n = n0;
  for (int nn = n0 - 1; nn <= n0 + 1; ++nn) {
    n = nn;
    if ((x / (r8)powi_f(10.f, nn)) >= 1.0 && (x / (r8)powi_f(10.f, nn)) < 10.0) break;
  }
  int idx = (int)(x / (r8)powi_f(10.f, n)) + 10 * (n - n2) - (n - n2);
  return std::max(1, std::min(idx, ntb));
}

}  // namespace

// ---- module state written by thompson_init, M:52, M:145, M:177, M:195-212, M:324-361 -----------
struct kor_handle {
  bool iiwarm, l_sediment, wp_double;
  r4 set_Nc, Nt_c, Sc3, D0i, xm0s, xm0g;
  r4 cce[6][16], ccg[6][16], ocg1[16], ocg2[16];
  r4 cie[8], cig[8], oig1, oig2, obmi;
  r4 cre[14], crg[14], ore1, org1, org2, org3, obmr;
  r4 cse[19], csg[19], oams, obms, ocms;
  r4 cge[13], cgg[13], oge1, ogg1, ogg2, ogg3, oamg, obmg, ocmg;
  r4 t1_qr_qc, t1_qr_qi, t2_qr_qi, t1_qg_qc, t1_qs_qc, t1_qs_qi, t1_qr_ev, t2_qr_ev;
  r4 t1_qs_sd, t2_qs_sd, t1_qg_sd, t2_qg_sd, t1_qs_me, t2_qs_me, t1_qg_me, t2_qg_me;
  int nic1, nic2, nii2, nii3, nir2, nir3, nis2, nig2, nig3, niIN2;
  r8 Dc[101], dtc[101], Di[101], dti[101], Dr[101], dtr[101], Ds[101], dts[101], Dg[101], dtg[101], t_Nc[101];
  r4 r_c[38], r_i[65], r_r[38], r_g[29], r_s[29], N0r_exp[38], N0g_exp[29], Nt_i[56], Nt_IN[56];
  // tables, column-major 1-based (accessors below)
  std::vector<r8> tcg_racg, tmr_racg, tcr_gacr, tmg_gacr, tnr_racg, tnr_gacr;
  std::vector<r8> tcs_racs1, tmr_racs1, tcs_racs2, tmr_racs2, tcr_sacr1, tms_sacr1, tcr_sacr2,
      tms_sacr2, tnr_racs1, tnr_racs2, tnr_sacr1, tnr_sacr2;
  std::vector<r8> tpi_qcfz, tni_qcfz, tpi_qrfz, tpg_qrfz, tni_qrfz, tnr_qrfz;
  std::vector<r8> tps_iaus, tni_iaus, tpi_ide, t_Efrw, t_Efsw;
  std::vector<r8> tpc_wev, tnc_wev;          // (nbc, ntb_c, nbc), table_dropEvap M:4400-4439 (read under is_aerosol_aware only)
  double init_seconds;
  int nthreads;
  std::map<std::string, std::vector<r8>*> tabs;
  std::map<std::string, std::vector<r8> > scal;
};

namespace {
typedef kor_handle H;
inline size_t ix4(int i, int j, int k, int m, int n1, int n2, int n3) {
  return (size_t)(i - 1) + (size_t)n1 * ((size_t)(j - 1) + (size_t)n2 * ((size_t)(k - 1) + (size_t)n3 * (size_t)(m - 1)));
}
inline size_t ix3(int i, int j, int k, int n1, int n2) { return (size_t)(i - 1) + (size_t)n1 * ((size_t)(j - 1) + (size_t)n2 * (size_t)(k - 1)); }
inline size_t ix2(int i, int j, int n1) { return (size_t)(i - 1) + (size_t)n1 * (size_t)(j - 1); }
#define RACG(t, i, j, k, m) (t)[ix4(i, j, k, m, ntb_g1, ntb_g, ntb_r1)]
#define RACS(t, i, j, k, m) (t)[ix4(i, j, k, m, ntb_s, ntb_t, ntb_r1)]
#define QRFZ(t, i, j, k) (t)[ix3(i, j, k, ntb_r, ntb_r1)]
#define QCFZ(t, i, k) (t)[ix2(i, k, ntb_c)]
#define IAUS(t, i, j) (t)[ix2(i, j, ntb_i)]
#define EFRW(t, i, j) (t)[ix2(i, j, nbr)]
#define EFSW(t, i, j) (t)[ix2(i, j, nbs)]

// M:604-670 size bins.  DFLOAT(x) = real(x,kind=wp) (M:8): f32 by default (U5).
void make_bins(const H& h, r8 lo, r8 hi, int nb, r8* D, r8* dt) {
  r8 xDx[102];
  xDx[1] = lo; xDx[nb + 1] = hi;
  for (int n = 2; n <= nb; ++n) {
    r8 frac = h.wp_double ? ((r8)(n - 1) / (r8)nb) : (r8)((r4)(n - 1) / (r4)nb);
    xDx[n] = exp(frac * log(xDx[nb + 1] / xDx[1]) + log(xDx[1]));
  }
  for (int n = 1; n <= nb; ++n) { D[n] = sqrt(xDx[n] * xDx[n + 1]); if (dt) dt[n] = xDx[n + 1] - xDx[n]; }
}

// rain fall speed polynomial used by the table builders, M:3733-3735, M:4278-4280
inline r8 vr_poly(r8 D) {
  return (r8)-0.1021f + (r8)4.932E3f * D - (r8)0.9551E6f * D * D + (r8)0.07934E9f * D * D * D
         - (r8)0.002362E12f * D * D * D * D;
}

// M:4243-4299
void table_Efrw(H& h) {
  for (int j = 1; j <= nbc; ++j)
    for (int i = 1; i <= nbr; ++i) {
      r8 Ef_rw = 0.0;
      r8 p = h.Dc[j] / h.Dr[i];
      if (h.Dr[i] < (r8)50.E-6f || h.Dc[j] < (r8)3.E-6f) { EFRW(h.t_Efrw, i, j) = 0.0; }
      else if (p > (r8)0.25f) {
        r8 X = h.Dc[j] * 1.E6;
        if (h.Dr[i] < (r8)75.e-6f) Ef_rw = (r8)0.026794f * X - (r8)0.20604f;
        else if (h.Dr[i] < (r8)125.e-6f) Ef_rw = (r8)-0.00066842f * X * X + (r8)0.061542f * X - (r8)0.37089f;
        else if (h.Dr[i] < (r8)175.e-6f)
          Ef_rw = (r8)4.091e-06f * X * X * X * X - (r8)0.00030908f * X * X * X + (r8)0.0066237f * X * X
                  - (r8)0.0013687f * X - (r8)0.073022f;
        else if (h.Dr[i] < (r8)250.e-6f)
          Ef_rw = (r8)9.6719e-5f * X * X * X - (r8)0.0068901f * X * X + (r8)0.17305f * X - (r8)0.65988f;
        else if (h.Dr[i] < (r8)350.e-6f)
          Ef_rw = (r8)9.0488e-5f * X * X * X - (r8)0.006585f * X * X + (r8)0.16606f * X - (r8)0.56125f;
        else
          Ef_rw = (r8)0.00010721f * X * X * X - (r8)0.0072962f * X * X + (r8)0.1704f * X - (r8)0.46929f;
      } else {
        r8 vtr = vr_poly(h.Dr[i]);
        r8 stokes = h.Dc[j] * h.Dc[j] * vtr * (r8)rho_w / ((r8)(9.f * 1.718E-5f) * h.Dr[i]);
        r8 reynolds = (r8)9.f * stokes / (p * p * (r8)rho_w);
        r8 F = log(reynolds);
        r8 G = -0.1007 - 0.358 * F + 0.0261 * F * F;
        r8 K0 = exp(G);
        r8 z = log(stokes / (K0 + 1.E-15));
        r8 Hh = 0.1465 + 1.302 * z - 0.607 * z * z + 0.293 * z * z * z;
        r8 yc0 = 2.0 / (r8)PI * atan(Hh);
        Ef_rw = (yc0 + p) * (yc0 + p) / (((r8)1.f + p) * ((r8)1.f + p));
      }
      // M:4294 runs for every branch (the zero branch leaves Ef_rw = 0)
      EFRW(h.t_Efrw, i, j) = (r8)maxf(0.0f, minf((r4)Ef_rw, 0.95f));
    }
}

// M:4307-4343
void table_Efsw(H& h) {
  for (int j = 1; j <= nbc; ++j) {
    r8 vtc = 1.19E4 * (1.0E4 * h.Dc[j] * h.Dc[j] * 0.25);
    for (int i = 1; i <= nbs; ++i) {
      r8 vts = (r8)av_s * pow(h.Ds[i], (r8)bv_s) * exp(-(r8)fv_s * h.Ds[i]) - vtc;
      r8 Ds_m = pow((r8)am_s * pow(h.Ds[i], (r8)bm_s) / (r8)am_r, (r8)h.obmr);
      r8 p = h.Dc[j] / Ds_m;
      if (p > (r8)0.25f || h.Ds[i] < (r8)D0s || h.Dc[j] < (r8)6.E-6f || vts < (r8)1.E-3f) {
        EFSW(h.t_Efsw, i, j) = 0.0;
      } else {
        r8 stokes = h.Dc[j] * h.Dc[j] * vts * (r8)rho_w / ((r8)(9.f * 1.718E-5f) * Ds_m);
        r8 reynolds = (r8)9.f * stokes / (p * p * (r8)rho_w);
        r8 F = log(reynolds);
        r8 G = -0.1007 - 0.358 * F + 0.0261 * F * F;
        r8 K0 = exp(G);
        r8 z = log(stokes / (K0 + 1.E-15));
        r8 Hh = 0.1465 + 1.302 * z - 0.607 * z * z + 0.293 * z * z * z;
        r8 yc0 = 2.0 / (r8)PI * atan(Hh);
        r8 Ef_sw = (yc0 + p) * (yc0 + p) / (((r8)1.f + p) * ((r8)1.f + p));
        EFSW(h.t_Efsw, i, j) = (r8)maxf(0.0f, minf((r4)Ef_sw, 0.95f));
      }
    }
  }
}

// rain PSD for table axis entry (m: r_r index, k: N0r_exp index), M:3755-3760
void rain_psd(const H& h, int m, int k, r8* N_r, r8* lamr_out, r8* N0r_out) {
  r8 lam_exp = (r8)powf(h.N0r_exp[k] * am_r * h.crg[1] / h.r_r[m], h.ore1);
  r8 lamr = lam_exp * (r8)powf(h.crg[3] * h.org2 * h.org1, h.obmr);
  r8 N0_r = (r8)h.N0r_exp[k] / ((r8)h.crg[2] * lam_exp) * pow(lamr, (r8)h.cre[2]);
  if (N_r)
    for (int n2 = 1; n2 <= nbr; ++n2)
      N_r[n2] = N0_r * pow(h.Dr[n2], (r8)mu_r) * exp(-lamr * h.Dr[n2]) * h.dtr[n2];
  if (lamr_out) *lamr_out = lamr;
  if (N0r_out) *N0r_out = N0_r;
}

// M:3698-3833 (file cache handled by the caller)
void qr_acr_qg(H& h) {
  r8 vg[101], vr[101];
  for (int n2 = 1; n2 <= nbr; ++n2) vr[n2] = vr_poly(h.Dr[n2]);
  for (int n = 1; n <= nbg; ++n) vg[n] = (r8)av_g * pow(h.Dg[n], (r8)bv_g);
  const r4 c0 = PI * .25f * Ef_rg;
  const int km_e = ntb_r * ntb_r1 - 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(h.nthreads)
  for (int km = 0; km <= km_e; ++km) {
    int m = km / ntb_r1 + 1;
    int k = km % ntb_r1 + 1;
    r8 N_r[101], N_g[101];
    rain_psd(h, m, k, N_r, 0, 0);
    for (int j = 1; j <= ntb_g; ++j)
      for (int i = 1; i <= ntb_g1; ++i) {
        r8 lam_exp = (r8)powf(h.N0g_exp[i] * am_g * h.cgg[1] / h.r_g[j], h.oge1);
        r8 lamg = lam_exp * (r8)powf(h.cgg[3] * h.ogg2 * h.ogg1, h.obmg);
        r8 N0_g = (r8)h.N0g_exp[i] / ((r8)h.cgg[2] * lam_exp) * pow(lamg, (r8)h.cge[2]);
        for (int n = 1; n <= nbg; ++n)
          N_g[n] = N0_g * pow(h.Dg[n], (r8)mu_g) * exp(-lamg * h.Dg[n]) * h.dtg[n];
        r8 t1 = 0, t2 = 0, z1 = 0, z2 = 0, y1 = 0, y2 = 0;
        for (int n2 = 1; n2 <= nbr; ++n2) {
          r8 massr = (r8)am_r * pow(h.Dr[n2], (r8)bm_r);
          for (int n = 1; n <= nbg; ++n) {
            r8 massg = (r8)am_g * pow(h.Dg[n], (r8)bm_g);
            r8 dvg = 0.5 * ((vr[n2] - vg[n]) + fabs(vr[n2] - vg[n]));
            r8 dvr = 0.5 * ((vg[n] - vr[n2]) + fabs(vg[n] - vr[n2]));
            r8 s = h.Dg[n] + h.Dr[n2];
            // adding an exact +0.0 term never changes a non-negative running sum, so the branch
            // on the sign of (vr-vg) is bit-identical to M:3785-3797 which adds both families.
            if (dvg != 0.0) {
              t1 = t1 + (r8)c0 * s * s * dvg * massg * N_g[n] * N_r[n2];
              z1 = z1 + (r8)c0 * s * s * dvg * massr * N_g[n] * N_r[n2];
              y1 = y1 + (r8)c0 * s * s * dvg * N_g[n] * N_r[n2];
            }
            if (dvr != 0.0) {
              t2 = t2 + (r8)c0 * s * s * dvr * massr * N_g[n] * N_r[n2];
              y2 = y2 + (r8)c0 * s * s * dvr * N_g[n] * N_r[n2];
              z2 = z2 + (r8)c0 * s * s * dvr * massg * N_g[n] * N_r[n2];
            }
          }
        }
        RACG(h.tcg_racg, i, j, k, m) = t1;
        RACG(h.tmr_racg, i, j, k, m) = mind(z1, (r8)h.r_r[m] * 1.0);
        RACG(h.tcr_gacr, i, j, k, m) = t2;
        RACG(h.tmg_gacr, i, j, k, m) = z2;
        RACG(h.tnr_racg, i, j, k, m) = y1;
        RACG(h.tnr_gacr, i, j, k, m) = y2;
      }
  }
}

// f32 polynomial of Field et al. for an arbitrary moment order c, M:1590-1599 pattern
inline void field_ab(r4 tc0, r4 c, r4& loga_, r4& b_) {
  loga_ = sa[1] + sa[2] * tc0 + sa[3] * c + sa[4] * tc0 * c + sa[5] * tc0 * tc0 + sa[6] * c * c
          + sa[7] * tc0 * tc0 * c + sa[8] * tc0 * c * c + sa[9] * tc0 * tc0 * tc0 + sa[10] * c * c * c;
  b_ = sb[1] + sb[2] * tc0 + sb[3] * c + sb[4] * tc0 * c + sb[5] * tc0 * tc0 + sb[6] * c * c
       + sb[7] * tc0 * tc0 * c + sb[8] * tc0 * c * c + sb[9] * tc0 * tc0 * tc0 + sb[10] * c * c * c;
}

// M:3842-4082
void qr_acr_qs(H& h) {
  r8 vr[101], vs[101];
  for (int n2 = 1; n2 <= nbr; ++n2) vr[n2] = vr_poly(h.Dr[n2]);
  for (int n = 1; n <= nbs; ++n)
    vs[n] = (r8)(1.5f * av_s) * pow(h.Ds[n], (r8)bv_s) * exp(-(r8)fv_s * h.Ds[n]);
  const r4 c0 = PI * .25f * Ef_rs;
  const int km_e = ntb_r * ntb_r1 - 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(h.nthreads)
  for (int km = 0; km <= km_e; ++km) {
    int m = km / ntb_r1 + 1;
    int k = km % ntb_r1 + 1;
    r8 N_r[101], N_s[101];
    rain_psd(h, m, k, N_r, 0, 0);
    for (int j = 1; j <= ntb_t; ++j)
      for (int i = 1; i <= ntb_s; ++i) {
        r8 M2 = (r8)(h.r_s[i] * h.oams) * 1.0;
        r8 second, loga_, a_, b_;
        r4 la, bb;
        if (bm_s > 2.0f - 1.E-3f && bm_s < 2.0f + 1.E-3f) {   // U8: kept as written (M:3938)
          field_ab(Tc[j], bm_s, la, bb);
          loga_ = la; b_ = bb;
          a_ = pow((r8)10.0f, loga_);
          second = pow(M2 / a_, (r8)1.f / b_);
        } else {
          second = M2;
        }
        field_ab(Tc[j], h.cse[1], la, bb);
        loga_ = la; b_ = bb;
        a_ = pow((r8)10.0f, loga_);
        r8 M3 = a_ * pow(second, b_);
        r8 oM3 = (r8)1.f / M3;
        r8 Mrat = M2 * (M2 * oM3) * (M2 * oM3) * (M2 * oM3);
        r8 M0 = pow(M2 * oM3, (r8)mu_s);
        r8 slam1 = M2 * oM3 * (r8)Lam0;
        r8 slam2 = M2 * oM3 * (r8)Lam1;
        for (int n = 1; n <= nbs; ++n)
          N_s[n] = Mrat * ((r8)Kap0 * exp(-slam1 * h.Ds[n])
                           + (r8)Kap1 * M0 * pow(h.Ds[n], (r8)mu_s) * exp(-slam2 * h.Ds[n])) * h.dts[n];
        r8 t1 = 0, t2 = 0, t3 = 0, t4 = 0, z1 = 0, z2 = 0, z3 = 0, z4 = 0, y1 = 0, y2 = 0, y3 = 0, y4 = 0;
        for (int n2 = 1; n2 <= nbr; ++n2) {
          r8 massr = (r8)am_r * pow(h.Dr[n2], (r8)bm_r);
          for (int n = 1; n <= nbs; ++n) {
            r8 masss = (r8)am_s * pow(h.Ds[n], (r8)bm_s);
            r8 dvs = 0.5 * ((vr[n2] - vs[n]) + fabs(vr[n2] - vs[n]));
            r8 dvr = 0.5 * ((vs[n] - vr[n2]) + fabs(vs[n] - vr[n2]));
            r8 s = h.Ds[n] + h.Dr[n2];
            bool big = massr > (r8)1.5f * masss;
            if (dvs != 0.0) {   // exact-zero terms skipped, bit-identical (see qr_acr_qg)
              if (big) {
                t1 = t1 + (r8)c0 * s * s * dvs * masss * N_s[n] * N_r[n2];
                z1 = z1 + (r8)c0 * s * s * dvs * massr * N_s[n] * N_r[n2];
                y1 = y1 + (r8)c0 * s * s * dvs * N_s[n] * N_r[n2];
              } else {
                t3 = t3 + (r8)c0 * s * s * dvs * masss * N_s[n] * N_r[n2];
                z3 = z3 + (r8)c0 * s * s * dvs * massr * N_s[n] * N_r[n2];
                y3 = y3 + (r8)c0 * s * s * dvs * N_s[n] * N_r[n2];
              }
            }
            if (dvr != 0.0) {
              if (big) {
                t2 = t2 + (r8)c0 * s * s * dvr * massr * N_s[n] * N_r[n2];
                y2 = y2 + (r8)c0 * s * s * dvr * N_s[n] * N_r[n2];
                z2 = z2 + (r8)c0 * s * s * dvr * masss * N_s[n] * N_r[n2];
              } else {
                t4 = t4 + (r8)c0 * s * s * dvr * massr * N_s[n] * N_r[n2];
                y4 = y4 + (r8)c0 * s * s * dvr * N_s[n] * N_r[n2];
                z4 = z4 + (r8)c0 * s * s * dvr * masss * N_s[n] * N_r[n2];
              }
            }
          }
        }
        RACS(h.tcs_racs1, i, j, k, m) = t1;
        RACS(h.tmr_racs1, i, j, k, m) = mind(z1, (r8)h.r_r[m] * 1.0);
        RACS(h.tcs_racs2, i, j, k, m) = t3;
        RACS(h.tmr_racs2, i, j, k, m) = z3;
        RACS(h.tcr_sacr1, i, j, k, m) = t2;
        RACS(h.tms_sacr1, i, j, k, m) = z2;
        RACS(h.tcr_sacr2, i, j, k, m) = t4;
        RACS(h.tms_sacr2, i, j, k, m) = z4;
        RACS(h.tnr_racs1, i, j, k, m) = y1;
        RACS(h.tnr_racs2, i, j, k, m) = y3;
        RACS(h.tnr_sacr1, i, j, k, m) = y2;
        RACS(h.tnr_sacr2, i, j, k, m) = y4;
      }
  }
}

// M:4092-4175.  U4: the m = 1..ntb_IN loop overwrites every entry; only m = ntb_IN survives.
void freezeH2O(H& h) {
  r8 massr[101], massc[101];
  r8 orho_w = (r8)(1.f / rho_w);
  for (int n2 = 1; n2 <= nbr; ++n2) massr[n2] = (r8)am_r * pow(h.Dr[n2], (r8)bm_r);
  for (int n = 1; n <= nbc; ++n) massc[n] = (r8)am_r * pow(h.Dc[n], (r8)bm_r);
  const int m = ntb_IN;
  r4 T_adjust = maxf(-3.0f, minf(3.0f - log10f(h.Nt_IN[m]), 3.0f));
  for (int k = 1; k <= 45; ++k) {
    r8 kk = h.wp_double ? (r8)k : (r8)(r4)k;
    r8 Texp = exp(kk - (r8)T_adjust * 1.0) - 1.0;
    for (int j = 1; j <= ntb_r1; ++j)
      for (int i = 1; i <= ntb_r; ++i) {
        r8 lamr, N0_r;
        rain_psd(h, i, j, 0, &lamr, &N0_r);
        r8 sum1 = 0, sum2 = 0, sumn1 = 0, sumn2 = 0;
        for (int n2 = nbr; n2 >= 1; --n2) {
          r8 N_r = N0_r * pow(h.Dr[n2], (r8)mu_r) * exp(-lamr * h.Dr[n2]) * h.dtr[n2];
          r8 vol = massr[n2] * orho_w;
          r8 prob = 1.0 - exp(-120.0 * vol * 5.2E-4 * Texp);
          if (massr[n2] < (r8)h.xm0g) {
            sumn1 = sumn1 + prob * N_r;
            sum1 = sum1 + prob * N_r * massr[n2];
          } else {
            sumn2 = sumn2 + prob * N_r;
            sum2 = sum2 + prob * N_r * massr[n2];
          }
        }
        QRFZ(h.tpi_qrfz, i, j, k) = sum1;
        QRFZ(h.tni_qrfz, i, j, k) = sumn1;
        QRFZ(h.tpg_qrfz, i, j, k) = sum2;
        QRFZ(h.tnr_qrfz, i, j, k) = sumn2;
      }
    int nu_c = std::min(15, nint_d((r8)1000.E6f / h.t_Nc[1]) + 2);
    for (int i = 1; i <= ntb_c; ++i) {
      r8 lamc = pow(h.t_Nc[1] * (r8)am_r * (r8)h.ccg[2][nu_c] * (r8)h.ocg1[nu_c] / (r8)h.r_c[i], (r8)h.obmr);
      r8 N0_c = h.t_Nc[1] * (r8)h.ocg1[nu_c] * pow(lamc, (r8)h.cce[1][nu_c]);
      r8 sum1 = 0, sumn2 = 0;
      for (int n = nbc; n >= 1; --n) {
        r8 vol = massc[n] * orho_w;
        r8 prob = 1.0 - exp(-120.0 * vol * 5.2E-4 * Texp);
        r8 N_c = N0_c * powi_d(h.Dc[n], nu_c) * exp(-lamc * h.Dc[n]) * h.dtc[n];
        sumn2 = mind(h.t_Nc[1], sumn2 + prob * N_c);
        sum1 = sum1 + prob * N_c * massc[n];
        if (sum1 >= (r8)h.r_c[i]) break;
      }
      QCFZ(h.tpi_qcfz, i, k) = sum1;
      QCFZ(h.tni_qcfz, i, k) = sumn2;
    }
  }
}

// M:4400-4439: mass and number of the cloud droplets smaller than D-star (they evaporate in one step)
#define WEV(t, i, j, k) t[((i) - 1) + (size_t)nbc * (((j) - 1) + (size_t)ntb_c * ((k) - 1))]
void table_dropEvap(H& h) {
  r8 N_c[nbc + 1], massc[nbc + 1];
  for (int n = 1; n <= nbc; ++n) massc[n] = (r8)am_r * pow(h.Dc[n], (r8)bm_r);
  for (int k = 1; k <= nbc; ++k) {
    const int nu_c = std::min(15, nint_d((r8)1000.E6f / h.t_Nc[k]) + 2);
    for (int j = 1; j <= ntb_c; ++j) {
      const r8 lamc = pow(h.t_Nc[k] * (r8)am_r * (r8)h.ccg[2][nu_c] * (r8)h.ocg1[nu_c] / (r8)h.r_c[j], (r8)h.obmr);
      const r8 N0_c = h.t_Nc[k] * (r8)h.ocg1[nu_c] * pow(lamc, (r8)h.cce[1][nu_c]);
      for (int i = 1; i <= nbc; ++i) {
        N_c[i] = N0_c * powi_d(h.Dc[i], nu_c) * exp(-lamc * h.Dc[i]) * h.dtc[i];
        r8 summ = 0., summ2 = 0.;
        for (int n = 1; n <= i; ++n) { summ = summ + massc[n] * N_c[n]; summ2 = summ2 + N_c[n]; }
        WEV(h.tpc_wev, i, j, k) = summ;
        WEV(h.tnc_wev, i, j, k) = summ2;
      }
    }
  }
}

// M:4190-4233
void qi_aut_qs(H& h) {
  for (int j = 1; j <= ntb_i1; ++j)
    for (int i = 1; i <= ntb_i; ++i) {
      r8 lami = (r8)powf(am_i * h.cig[2] * h.oig1 * h.Nt_i[j] / h.r_i[i], h.obmi);
      r8 Di_mean = (r8)(bm_i + mu_i + 1.f) / lami;
      r8 N0_i = (r8)(h.Nt_i[j] * h.oig1) * pow(lami, (r8)h.cie[1]);
      r8 t1 = 0, t2 = 0;
      if ((r4)Di_mean > 5.f * D0s) {
        t1 = h.r_i[i]; t2 = h.Nt_i[j];
        IAUS(h.tpi_ide, i, j) = 0.0;
      } else if ((r4)Di_mean < h.D0i) {
        t1 = 0; t2 = 0;
        IAUS(h.tpi_ide, i, j) = 1.0;
      } else {
        r4 xlimit_intg = (r4)(lami * (r8)D0s);
        IAUS(h.tpi_ide, i, j) = (r8)GAMMP(mu_i + 2.0f, xlimit_intg) * 1.0;
        for (int n2 = 1; n2 <= nbi; ++n2) {
          r8 N_i = N0_i * pow(h.Di[n2], (r8)mu_i) * exp(-lami * h.Di[n2]) * h.dti[n2];
          if (h.Di[n2] >= (r8)D0s) {
            t1 = t1 + N_i * (r8)am_i * pow(h.Di[n2], (r8)bm_i);
            t2 = t2 + N_i;
          }
        }
      }
      IAUS(h.tps_iaus, i, j) = t1;
      IAUS(h.tni_iaus, i, j) = t2;
    }
}

uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* c = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
  return h;
}

// M:374-797
void thompson_init(H& h, const char* cache_path) {
  auto t0 = std::chrono::steady_clock::now();
  h.Nt_c = h.set_Nc * 1.e6f;
  make_axis(h.r_c, ntb_c, -6, -2);
  make_axis(h.r_i, ntb_i, -10, -3);
  make_axis(h.r_r, ntb_r, -6, -2);
  make_axis(h.r_g, ntb_g, -5, -2);
  make_axis(h.r_s, ntb_s, -5, -2);
  make_axis(h.N0r_exp, ntb_r1, 6, 10);
  make_axis(h.N0g_exp, ntb_g1, 4, 7);
  make_axis(h.Nt_i, ntb_i1, 0, 6);
  make_axis(h.Nt_IN, ntb_IN, 0, 6);

  h.Sc3 = powf(Sc, 1.f / 3.f);
  h.D0i = powf(xm0i / am_i, 1.f / bm_i);
  h.xm0s = am_s * powf(D0s, bm_s);
  h.xm0g = am_g * powf(D0g, bm_g);

  for (int n = 1; n <= 15; ++n) {
    h.cce[1][n] = (r4)n + 1.f;
    h.cce[2][n] = bm_r + (r4)n + 1.f;
    h.cce[3][n] = bm_r + (r4)n + 4.f;
    h.cce[4][n] = (r4)n + bv_c + 1.f;
    h.cce[5][n] = bm_r + (r4)n + bv_c + 1.f;
    for (int q = 1; q <= 5; ++q) h.ccg[q][n] = WGAMMA(h.cce[q][n]);
    h.ocg1[n] = 1.f / h.ccg[1][n];
    h.ocg2[n] = 1.f / h.ccg[2][n];
  }
  h.cie[1] = mu_i + 1.f;
  h.cie[2] = bm_i + mu_i + 1.f;
  h.cie[3] = bm_i + mu_i + bv_i + 1.f;
  h.cie[4] = mu_i + bv_i + 1.f;
  h.cie[5] = mu_i + 2.f;
  h.cie[6] = bm_i * 0.5f + mu_i + bv_i + 1.f;
  h.cie[7] = bm_i * 0.5f + mu_i + 1.f;
  for (int n = 1; n <= 7; ++n) h.cig[n] = WGAMMA(h.cie[n]);
  h.oig1 = 1.f / h.cig[1]; h.oig2 = 1.f / h.cig[2]; h.obmi = 1.f / bm_i;

  h.cre[1] = bm_r + 1.f;
  h.cre[2] = mu_r + 1.f;
  h.cre[3] = bm_r + mu_r + 1.f;
  h.cre[4] = bm_r * 2.f + mu_r + 1.f;
  h.cre[5] = mu_r + bv_r + 1.f;
  h.cre[6] = bm_r + mu_r + bv_r + 1.f;
  h.cre[7] = bm_r * 0.5f + mu_r + bv_r + 1.f;
  h.cre[8] = bm_r + mu_r + bv_r + 3.f;
  h.cre[9] = mu_r + bv_r + 3.f;
  h.cre[10] = mu_r + 2.f;
  h.cre[11] = 0.5f * (bv_r + 5.f + 2.f * mu_r);
  h.cre[12] = bm_r * 0.5f + mu_r + 1.f;
  h.cre[13] = bm_r * 2.f + mu_r + bv_r + 1.f;
  for (int n = 1; n <= 13; ++n) h.crg[n] = WGAMMA(h.cre[n]);
  h.obmr = 1.f / bm_r; h.ore1 = 1.f / h.cre[1];
  h.org1 = 1.f / h.crg[1]; h.org2 = 1.f / h.crg[2]; h.org3 = 1.f / h.crg[3];

  h.cse[1] = bm_s + 1.f;
  h.cse[2] = bm_s + 2.f;
  h.cse[3] = bm_s * 2.f;
  h.cse[4] = bm_s + bv_s + 1.f;
  h.cse[5] = bm_s * 2.f + bv_s + 1.f;
  h.cse[6] = bm_s * 2.f + 1.f;
  h.cse[7] = bm_s + mu_s + 1.f;
  h.cse[8] = bm_s + mu_s + 2.f;
  h.cse[9] = bm_s + mu_s + 3.f;
  h.cse[10] = bm_s + mu_s + bv_s + 1.f;
  h.cse[11] = bm_s * 2.f + mu_s + bv_s + 1.f;
  h.cse[12] = bm_s * 2.f + mu_s + 1.f;
  h.cse[13] = bv_s + 2.f;
  h.cse[14] = bm_s + bv_s;
  h.cse[15] = mu_s + 1.f;
  h.cse[16] = 1.0f + (1.0f + bv_s) / 2.f;
  h.cse[17] = h.cse[16] + mu_s + 1.f;
  h.cse[18] = bv_s + mu_s + 3.f;
  for (int n = 1; n <= 18; ++n) h.csg[n] = WGAMMA(h.cse[n]);
  h.oams = 1.f / am_s; h.obms = 1.f / bm_s; h.ocms = powf(h.oams, h.obms);

  h.cge[1] = bm_g + 1.f;
  h.cge[2] = mu_g + 1.f;
  h.cge[3] = bm_g + mu_g + 1.f;
  h.cge[4] = bm_g * 2.f + mu_g + 1.f;
  h.cge[5] = bm_g * 2.f + mu_g + bv_g + 1.f;
  h.cge[6] = bm_g + mu_g + bv_g + 1.f;
  h.cge[7] = bm_g + mu_g + bv_g + 2.f;
  h.cge[8] = bm_g + mu_g + bv_g + 3.f;
  h.cge[9] = mu_g + bv_g + 3.f;
  h.cge[10] = mu_g + 2.f;
  h.cge[11] = 0.5f * (bv_g + 5.f + 2.f * mu_g);
  h.cge[12] = 0.5f * (bv_g + 5.f) + mu_g;
  for (int n = 1; n <= 12; ++n) h.cgg[n] = WGAMMA(h.cge[n]);
  h.oamg = 1.f / am_g; h.obmg = 1.f / bm_g; h.ocmg = powf(h.oamg, h.obmg);
  h.oge1 = 1.f / h.cge[1];
  h.ogg1 = 1.f / h.cgg[1]; h.ogg2 = 1.f / h.cgg[2]; h.ogg3 = 1.f / h.cgg[3];

  // M:559-591
  h.t1_qr_qc = PI * .25f * av_r * h.crg[9];
  h.t1_qr_qi = PI * .25f * av_r * h.crg[9];
  h.t2_qr_qi = PI * .25f * am_r * av_r * h.crg[8];
  h.t1_qg_qc = PI * .25f * av_g * h.cgg[9];
  h.t1_qs_qc = PI * .25f * av_s;
  h.t1_qs_qi = PI * .25f * av_s;
  h.t1_qr_ev = 0.78f * h.crg[10];
  h.t2_qr_ev = 0.308f * h.Sc3 * sqrtf(av_r) * h.crg[11];
  h.t1_qs_sd = 0.86f;
  h.t2_qs_sd = 0.28f * h.Sc3 * sqrtf(av_s);
  h.t1_qs_me = PI * 4.f * C_sqrd * olfus * 0.86f;
  h.t2_qs_me = PI * 4.f * C_sqrd * olfus * 0.28f * h.Sc3 * sqrtf(av_s);
  h.t1_qg_sd = 0.86f * h.cgg[10];
  h.t2_qg_sd = 0.28f * h.Sc3 * sqrtf(av_g) * h.cgg[11];
  h.t1_qg_me = PI * 4.f * C_cube * olfus * 0.86f * h.cgg[10];
  h.t2_qg_me = PI * 4.f * C_cube * olfus * 0.28f * h.Sc3 * sqrtf(av_g) * h.cgg[11];

  // M:594-602
  h.nic2 = nint_f(log10f(h.r_c[1]));
  h.nii2 = nint_f(log10f(h.r_i[1]));
  h.nii3 = nint_f(log10f(h.Nt_i[1]));
  h.nir2 = nint_f(log10f(h.r_r[1]));
  h.nir3 = nint_f(log10f(h.N0r_exp[1]));
  h.nis2 = nint_f(log10f(h.r_s[1]));
  h.nig2 = nint_f(log10f(h.r_g[1]));
  h.nig3 = nint_f(log10f(h.N0g_exp[1]));
  h.niIN2 = nint_f(log10f(h.Nt_IN[1]));

  // M:604-670
  h.Dc[1] = (r8)D0c * 1.0; h.dtc[1] = (r8)D0c * 1.0;
  for (int n = 2; n <= nbc; ++n) { h.Dc[n] = h.Dc[n - 1] + 1.0E-6; h.dtc[n] = h.Dc[n] - h.Dc[n - 1]; }
  make_bins(h, (r8)h.D0i * 1.0, 5.0 * (r8)D0s, nbi, h.Di, h.dti);
  make_bins(h, (r8)D0r * 1.0, 0.005, nbr, h.Dr, h.dtr);
  make_bins(h, (r8)D0s * 1.0, 0.02, nbs, h.Ds, h.dts);
  make_bins(h, (r8)D0g * 1.0, 0.05, nbg, h.Dg, h.dtg);
  make_bins(h, 1.0, 3000.0, nbc, h.t_Nc, 0);
  for (int n = 1; n <= nbc; ++n) h.t_Nc[n] *= 1.E6;
  h.nic1 = (int)log(h.t_Nc[nbc] / h.t_Nc[1]);

  // allocate + zero, M:386-430, M:676-750
  size_t n_racg = (size_t)ntb_g1 * ntb_g * ntb_r1 * ntb_r, n_racs = (size_t)ntb_s * ntb_t * ntb_r1 * ntb_r;
  std::vector<r8>* g6[] = {&h.tcg_racg, &h.tmr_racg, &h.tcr_gacr, &h.tmg_gacr, &h.tnr_racg, &h.tnr_gacr};
  const char* g6n[] = {"tcg_racg", "tmr_racg", "tcr_gacr", "tmg_gacr", "tnr_racg", "tnr_gacr"};
  for (int q = 0; q < 6; ++q) { g6[q]->assign(n_racg, 0.0); h.tabs[g6n[q]] = g6[q]; }
  std::vector<r8>* s12[] = {&h.tcs_racs1, &h.tmr_racs1, &h.tcs_racs2, &h.tmr_racs2, &h.tcr_sacr1, &h.tms_sacr1,
                            &h.tcr_sacr2, &h.tms_sacr2, &h.tnr_racs1, &h.tnr_racs2, &h.tnr_sacr1, &h.tnr_sacr2};
  const char* s12n[] = {"tcs_racs1", "tmr_racs1", "tcs_racs2", "tmr_racs2", "tcr_sacr1", "tms_sacr1",
                        "tcr_sacr2", "tms_sacr2", "tnr_racs1", "tnr_racs2", "tnr_sacr1", "tnr_sacr2"};
  for (int q = 0; q < 12; ++q) { s12[q]->assign(n_racs, 0.0); h.tabs[s12n[q]] = s12[q]; }
  h.tpi_qcfz.assign(ntb_c * 45, 0.0); h.tni_qcfz.assign(ntb_c * 45, 0.0);
  h.tpi_qrfz.assign(ntb_r * ntb_r1 * 45, 0.0); h.tpg_qrfz.assign(ntb_r * ntb_r1 * 45, 0.0);
  h.tni_qrfz.assign(ntb_r * ntb_r1 * 45, 0.0); h.tnr_qrfz.assign(ntb_r * ntb_r1 * 45, 0.0);
  h.tps_iaus.assign(ntb_i * ntb_i1, 0.0); h.tni_iaus.assign(ntb_i * ntb_i1, 0.0); h.tpi_ide.assign(ntb_i * ntb_i1, 0.0);
  h.t_Efrw.assign(nbr * nbc, 0.0); h.t_Efsw.assign(nbs * nbc, 0.0);
  h.tabs["tpi_qcfz"] = &h.tpi_qcfz; h.tabs["tni_qcfz"] = &h.tni_qcfz;
  h.tabs["tpi_qrfz"] = &h.tpi_qrfz; h.tabs["tpg_qrfz"] = &h.tpg_qrfz;
  h.tabs["tni_qrfz"] = &h.tni_qrfz; h.tabs["tnr_qrfz"] = &h.tnr_qrfz;
  h.tabs["tps_iaus"] = &h.tps_iaus; h.tabs["tni_iaus"] = &h.tni_iaus; h.tabs["tpi_ide"] = &h.tpi_ide;
  h.tabs["t_Efrw"] = &h.t_Efrw; h.tabs["t_Efsw"] = &h.t_Efsw;

  table_Efrw(h);
  table_Efsw(h);
  h.tpc_wev.assign((size_t)nbc * ntb_c * nbc, 0.0); h.tnc_wev.assign((size_t)nbc * ntb_c * nbc, 0.0);
  h.tabs["tpc_wev"] = &h.tpc_wev; h.tabs["tnc_wev"] = &h.tnc_wev;
  table_dropEvap(h);                                   // M:771 (only read under is_aerosol_aware, M:2804, M:2850)
  if (!h.iiwarm) {
    // binary cache of the two 4-D families (stand-in for run_data/*.data, M:3717-3728): keyed by
    // a hash of everything the builders read.
    uint64_t key = fnv(h.Dr + 1, 800); key = fnv(h.dtr + 1, 800, key); key = fnv(h.Dg + 1, 800, key);
    key = fnv(h.dtg + 1, 800, key); key = fnv(h.Ds + 1, 800, key); key = fnv(h.dts + 1, 800, key);
    key = fnv(h.crg, sizeof h.crg, key); key = fnv(h.cgg, sizeof h.cgg, key); key = fnv(h.cse, sizeof h.cse, key);
    key = fnv(h.cre, sizeof h.cre, key); key = fnv(h.cge, sizeof h.cge, key);
    key ^= 0x6b6f7232ull;  // format tag "kor2"
    bool loaded = false;
    if (cache_path && *cache_path) {
      FILE* f = fopen(cache_path, "rb");
      if (f) {
        uint64_t k2 = 0;
        if (fread(&k2, 8, 1, f) == 1 && k2 == key) {
          bool ok = true;
          for (int q = 0; q < 6 && ok; ++q) ok = fread(g6[q]->data(), 8, n_racg, f) == n_racg;
          for (int q = 0; q < 12 && ok; ++q) ok = fread(s12[q]->data(), 8, n_racs, f) == n_racs;
          loaded = ok;
        }
        fclose(f);
      }
    }
    if (!loaded) {
      qr_acr_qg(h);
      qr_acr_qs(h);
      if (cache_path && *cache_path) {
        std::string tmp = std::string(cache_path) + ".tmp" + std::to_string((long)getpid());
        FILE* f = fopen(tmp.c_str(), "wb");
        if (f) {
          fwrite(&key, 8, 1, f);
          for (int q = 0; q < 6; ++q) fwrite(g6[q]->data(), 8, n_racg, f);
          for (int q = 0; q < 12; ++q) fwrite(s12[q]->data(), 8, n_racs, f);
          fclose(f);
          rename(tmp.c_str(), cache_path);
        }
      }
    }
    freezeH2O(h);
    qi_aut_qs(h);
  }

  // scalars/vectors exposed for known-answer tests
  auto put = [&](const char* n, const r4* a, int cnt) { std::vector<r8> v(a, a + cnt); h.scal[n] = v; };
  auto putd = [&](const char* n, const r8* a, int cnt) { std::vector<r8> v(a, a + cnt); h.scal[n] = v; };
  put("cre", h.cre + 1, 13); put("crg", h.crg + 1, 13); put("cse", h.cse + 1, 18); put("csg", h.csg + 1, 18);
  put("cge", h.cge + 1, 12); put("cgg", h.cgg + 1, 12); put("cie", h.cie + 1, 7); put("cig", h.cig + 1, 7);
  for (int q = 1; q <= 5; ++q) { put(("cce" + std::to_string(q)).c_str(), h.cce[q] + 1, 15); put(("ccg" + std::to_string(q)).c_str(), h.ccg[q] + 1, 15); }
  put("ocg1", h.ocg1 + 1, 15); put("ocg2", h.ocg2 + 1, 15);
  r4 sc[] = {h.Nt_c, h.Sc3, h.D0i, h.xm0s, h.xm0g, rho_not, h.t1_qr_qc, h.t1_qr_qi, h.t2_qr_qi, h.t1_qg_qc, h.t1_qs_qc,
             h.t1_qs_qi, h.t1_qr_ev, h.t2_qr_ev, h.t1_qs_sd, h.t2_qs_sd, h.t1_qg_sd, h.t2_qg_sd, h.t1_qs_me,
             h.t2_qs_me, h.t1_qg_me, h.t2_qg_me, h.oig1, h.oig2, h.obmi, h.ore1, h.org1, h.org2, h.org3, h.obmr,
             h.oams, h.obms, h.ocms, h.oge1, h.ogg1, h.ogg2, h.ogg3, h.oamg, h.obmg, h.ocmg, am_r, am_g, am_i};
  put("scalars", sc, sizeof sc / sizeof sc[0]);
  r4 ii[] = {(r4)h.nic1, (r4)h.nic2, (r4)h.nii2, (r4)h.nii3, (r4)h.nir2, (r4)h.nir3, (r4)h.nis2, (r4)h.nig2, (r4)h.nig3, (r4)h.niIN2};
  put("offsets", ii, 10);
  putd("Dc", h.Dc + 1, 100); putd("Di", h.Di + 1, 100); putd("Dr", h.Dr + 1, 100); putd("Ds", h.Ds + 1, 100);
  putd("Dg", h.Dg + 1, 100); putd("t_Nc", h.t_Nc + 1, 100);
  putd("dtc", h.dtc + 1, 100); putd("dti", h.dti + 1, 100); putd("dtr", h.dtr + 1, 100); putd("dts", h.dts + 1, 100); putd("dtg", h.dtg + 1, 100);
  h.init_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

#include "thompson_oracle_step.inc"
