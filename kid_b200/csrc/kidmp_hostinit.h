// kidmp_hostinit.h - host half of thompson_init (M:374-797): the scalar constants, the size-bin
// grids and the per-axis-node scalars of the lookup tables.  A few thousand libm calls; the
// 1.4e10-term bin integrals that make the reference's init slow run on the device
// (kidmp_tables.cuh).  Host code only (compiled by nvcc's host compiler, no FMA contraction).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kidmp_internal.h"

namespace kidmp {
namespace hostinit {

// ---- Numerical-Recipes gamma family as the reference carries it (M:4530-4651) -----------------
inline float gammln(float xx) {                       // M:4598-4620: f64 inside, REAL result
  static const double cof[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                                -1.231739572450155, .1208650973866179E-2, -.5395239384953E-5};
  const double x = xx;
  double y = x, tmp = x + 5.5, ser = 1.000000000190015;
  tmp = (x + 0.5) * std::log(tmp) - tmp;
  for (int j = 0; j < 6; ++j) { y += 1.0; ser += cof[j] / y; }
  return (float)(tmp + std::log(2.5066282746310005 * ser / x));
}
inline float wgamma(float y) { return std::exp(gammln(y)); }          // M:4644-4651 (f32 EXP)
inline float gammp(float a, float x) {                                 // M:4623-4641 with GSER / GCF
  if (x < 0.f || a <= 0.f) return 0.f;
  const float gln = gammln(a);
  if (x < a + 1.f) {                                                   // series, M:4566-4594
    if (x <= 0.f) return 0.f;
    float ap = a, sum = 1.f / a, del = sum;
    for (int n = 1; n <= 100; ++n) {
      ap += 1.f; del = del * x / ap; sum += del;
      if (std::fabs(del) < std::fabs(sum) * 3.E-7f) break;
    }
    return sum * std::exp(-x + a * std::log(x) - gln);
  }
  const float fpmin = 1.E-30f;                                         // continued fraction, M:4530-4563
  float b = x + 1.f - a, c = 1.f / fpmin, d = 1.f / b, h = d;
  for (int i = 1; i <= 100; ++i) {
    const float an = -((float)i * ((float)i - a));
    b += 2.f;
    d = an * d + b; if (std::fabs(d) < fpmin) d = fpmin;
    c = b + an / c; if (std::fabs(c) < fpmin) c = fpmin;
    d = 1.f / d;
    const float del = d * c;
    h *= del;
    if (std::fabs(del - 1.f) < 3.E-7f) break;
  }
  return 1.f - std::exp(-x + a * std::log(x) - gln) * h;
}

// 10.**n the way gfortran evaluates real**integer (libgcc __powisf2: square-and-multiply)
inline float powi10(int m) {
  unsigned n = m < 0 ? (unsigned)(-m) : (unsigned)m;
  float x = 10.f, y = (n % 2) ? x : 1.0f;
  while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
  return m < 0 ? 1.0f / y : y;
}

// table axes M:215-303: mantissas 1..9 over decades [lo, hi), closed by 1.e<hi>; decimal
// literals are converted exactly like the compiler converts "3.e-5"
inline void axis(float* a, int n, int lo, int hi) {
  int c = 0;
  char buf[32];
  for (int d = lo; d < hi; ++d)
    for (int m = 1; m <= 9; ++m) { snprintf(buf, sizeof buf, "%d.e%d", m, d); a[c++] = strtof(buf, nullptr); }
  snprintf(buf, sizeof buf, "1.e%d", hi); a[c++] = strtof(buf, nullptr);
  if (c != n) abort();
}

// M:604-670 logarithmic size bins; DFLOAT() is real(.,kind=wp) (M:8), f32 unless wp_double (U5)
inline void bins(bool wp_double, double lo, double hi, double* D, double* dt) {
  double x[NBINS + 1];
  x[0] = lo; x[NBINS] = hi;
  for (int n = 1; n < NBINS; ++n) {
    const double frac = wp_double ? ((double)n / (double)NBINS) : (double)((float)n / (float)NBINS);
    x[n] = std::exp(frac * std::log(x[NBINS] / x[0]) + std::log(x[0]));
  }
  for (int n = 0; n < NBINS; ++n) { D[n] = std::sqrt(x[n] * x[n + 1]); if (dt) dt[n] = x[n + 1] - x[n]; }
}

struct Prep {   // per-axis-node scalars (host copies; see TablePrep in kidmp_tables.cuh)
  std::vector<double> lamr, N0_r, lamg, N0_g, s_Mrat, s_M0, s_slam1, s_slam2, lamc, N0_c, i_lami, i_N0, i_tpi_ide;
  std::vector<int> i_branch;
  double Texp[NTB_TC];
  int nu_c_fz;
  float am_s;
};

inline void compute(const kidmp_config& cfg, KConst& kc, HostBins& hb, Prep& pp) {
  const float PI = KP_PI;
  const float mu_s = KP_MU_S, bm_r = 3.f, bm_s = 2.f, bm_g = 3.f, bm_i = 3.f, bv_r = 1.f, bv_s = KP_BV_S,
              bv_g = KP_BV_G, bv_i = 1.f, bv_c = 2.f, mu_r = 0.f, mu_g = 0.f, mu_i = 0.f;
  const float am_r = PI * KP_RHO_W / 6.0f, am_g = PI * KP_RHO_G / 6.0f, am_i = PI * KP_RHO_I / 6.0f, am_s = KP_AM_S;
  const float Sc = 0.632f, Rv = 461.5f;
  const bool wpd = cfg.wp_double != 0;
  kc.Nt_c = cfg.set_Nc * 1.e6f;                                          // M:381
  kc.iiwarm = cfg.iiwarm != 0; kc.l_sediment = cfg.l_sediment != 0;
  kc.am_r = am_r; kc.am_g = am_g; kc.am_i = am_i;
  kc.oRv = 1.f / Rv; kc.lfus = KP_LSUB - KP_LVAP0; kc.olfus = 1.f / kc.lfus;
  kc.rho_not = 101325.0f / (287.05f * 298.0f);                           // M:141
  kc.Sc3 = std::pow(Sc, 1.f / 3.f);                                       // M:442-447
  kc.D0i = std::pow(KP_XM0I / am_i, 1.f / bm_i);
  kc.xm0s = am_s * std::pow(KP_D0S, bm_s);
  kc.xm0g = am_g * std::pow(KP_D0G, bm_g);

  for (int n = 1; n <= 15; ++n) {                                         // M:452-464
    const float fn = (float)n;
    kc.cce[0][n - 1] = fn + 1.f;
    kc.cce[1][n - 1] = bm_r + fn + 1.f;
    kc.cce[2][n - 1] = bm_r + fn + 4.f;
    kc.cce[3][n - 1] = fn + bv_c + 1.f;
    kc.cce[4][n - 1] = bm_r + fn + bv_c + 1.f;
    for (int q = 0; q < 5; ++q) kc.ccg[q][n - 1] = wgamma(kc.cce[q][n - 1]);
    kc.ocg1[n - 1] = 1.f / kc.ccg[0][n - 1];
    kc.ocg2[n - 1] = 1.f / kc.ccg[1][n - 1];
  }
  {                                                                        // M:467-483
    float* e = kc.cie;
    e[0] = mu_i + 1.f; e[1] = bm_i + mu_i + 1.f; e[2] = bm_i + mu_i + bv_i + 1.f; e[3] = mu_i + bv_i + 1.f;
    e[4] = mu_i + 2.f; e[5] = bm_i * 0.5f + mu_i + bv_i + 1.f; e[6] = bm_i * 0.5f + mu_i + 1.f;
    for (int n = 0; n < 7; ++n) kc.cig[n] = wgamma(e[n]);
    kc.oig1 = 1.f / kc.cig[0]; kc.oig2 = 1.f / kc.cig[1]; kc.obmi = 1.f / bm_i;
  }
  {                                                                        // M:485-504
    float* e = kc.cre;
    e[0] = bm_r + 1.f; e[1] = mu_r + 1.f; e[2] = bm_r + mu_r + 1.f; e[3] = bm_r * 2.f + mu_r + 1.f;
    e[4] = mu_r + bv_r + 1.f; e[5] = bm_r + mu_r + bv_r + 1.f; e[6] = bm_r * 0.5f + mu_r + bv_r + 1.f;
    e[7] = bm_r + mu_r + bv_r + 3.f; e[8] = mu_r + bv_r + 3.f; e[9] = mu_r + 2.f;
    e[10] = 0.5f * (bv_r + 5.f + 2.f * mu_r); e[11] = bm_r * 0.5f + mu_r + 1.f; e[12] = bm_r * 2.f + mu_r + bv_r + 1.f;
    for (int n = 0; n < 13; ++n) kc.crg[n] = wgamma(e[n]);
    kc.obmr = 1.f / bm_r; kc.ore1 = 1.f / e[0];
    kc.org1 = 1.f / kc.crg[0]; kc.org2 = 1.f / kc.crg[1]; kc.org3 = 1.f / kc.crg[2];
  }
  {                                                                        // M:507-530
    float* e = kc.cse;
    e[0] = bm_s + 1.f; e[1] = bm_s + 2.f; e[2] = bm_s * 2.f; e[3] = bm_s + bv_s + 1.f; e[4] = bm_s * 2.f + bv_s + 1.f;
    e[5] = bm_s * 2.f + 1.f; e[6] = bm_s + mu_s + 1.f; e[7] = bm_s + mu_s + 2.f; e[8] = bm_s + mu_s + 3.f;
    e[9] = bm_s + mu_s + bv_s + 1.f; e[10] = bm_s * 2.f + mu_s + bv_s + 1.f; e[11] = bm_s * 2.f + mu_s + 1.f;
    e[12] = bv_s + 2.f; e[13] = bm_s + bv_s; e[14] = mu_s + 1.f; e[15] = 1.0f + (1.0f + bv_s) / 2.f;
    e[16] = e[15] + mu_s + 1.f; e[17] = bv_s + mu_s + 3.f;
    for (int n = 0; n < 18; ++n) kc.csg[n] = wgamma(e[n]);
    kc.oams = 1.f / am_s; kc.obms = 1.f / bm_s; kc.ocms = std::pow(kc.oams, kc.obms);
  }
  {                                                                        // M:532-553
    float* e = kc.cge;
    e[0] = bm_g + 1.f; e[1] = mu_g + 1.f; e[2] = bm_g + mu_g + 1.f; e[3] = bm_g * 2.f + mu_g + 1.f;
    e[4] = bm_g * 2.f + mu_g + bv_g + 1.f; e[5] = bm_g + mu_g + bv_g + 1.f; e[6] = bm_g + mu_g + bv_g + 2.f;
    e[7] = bm_g + mu_g + bv_g + 3.f; e[8] = mu_g + bv_g + 3.f; e[9] = mu_g + 2.f;
    e[10] = 0.5f * (bv_g + 5.f + 2.f * mu_g); e[11] = 0.5f * (bv_g + 5.f) + mu_g;
    for (int n = 0; n < 12; ++n) kc.cgg[n] = wgamma(e[n]);
    kc.oamg = 1.f / am_g; kc.obmg = 1.f / bm_g; kc.ocmg = std::pow(kc.oamg, kc.obmg);
    kc.oge1 = 1.f / e[0];
    kc.ogg1 = 1.f / kc.cgg[0]; kc.ogg2 = 1.f / kc.cgg[1]; kc.ogg3 = 1.f / kc.cgg[2];
  }
  // rate prefactors, M:559-591
  kc.t1_qr_qc = PI * .25f * KP_AV_R * kc.crg[8];
  kc.t1_qr_qi = PI * .25f * KP_AV_R * kc.crg[8];
  kc.t2_qr_qi = PI * .25f * am_r * KP_AV_R * kc.crg[7];
  kc.t1_qg_qc = PI * .25f * KP_AV_G * kc.cgg[8];
  kc.t1_qs_qc = PI * .25f * KP_AV_S;
  kc.t1_qs_qi = PI * .25f * KP_AV_S;
  kc.t1_qr_ev = 0.78f * kc.crg[9];
  kc.t2_qr_ev = 0.308f * kc.Sc3 * std::sqrt(KP_AV_R) * kc.crg[10];
  kc.t1_qs_sd = 0.86f;
  kc.t2_qs_sd = 0.28f * kc.Sc3 * std::sqrt(KP_AV_S);
  kc.t1_qs_me = PI * 4.f * KP_C_SQRD * kc.olfus * 0.86f;
  kc.t2_qs_me = PI * 4.f * KP_C_SQRD * kc.olfus * 0.28f * kc.Sc3 * std::sqrt(KP_AV_S);
  kc.t1_qg_sd = 0.86f * kc.cgg[9];
  kc.t2_qg_sd = 0.28f * kc.Sc3 * std::sqrt(KP_AV_G) * kc.cgg[10];
  kc.t1_qg_me = PI * 4.f * KP_C_CUBE * kc.olfus * 0.86f * kc.cgg[9];
  kc.t2_qg_me = PI * 4.f * KP_C_CUBE * kc.olfus * 0.28f * kc.Sc3 * std::sqrt(KP_AV_G) * kc.cgg[10];

  // axes and decade offsets, M:215-303, M:594-602
  float Nt_IN[NTB_I1];
  axis(hb.r_c, NTB_C, -6, -2); axis(hb.r_i, NTB_I, -10, -3); axis(hb.r_r, NTB_R, -6, -2);
  axis(hb.r_g, NTB_G, -5, -2); axis(hb.r_s, NTB_S, -5, -2); axis(hb.N0r_exp, NTB_R1, 6, 10);
  axis(hb.N0g_exp, NTB_G1, 4, 7); axis(hb.Nt_i, NTB_I1, 0, 6); axis(Nt_IN, NTB_I1, 0, 6);
  auto dec = [](float v) { return (int)std::lround(std::log10(v)); };
  kc.nic2 = dec(hb.r_c[0]); kc.nii2 = dec(hb.r_i[0]); kc.nii3 = dec(hb.Nt_i[0]); kc.nir2 = dec(hb.r_r[0]);
  kc.nir3 = dec(hb.N0r_exp[0]); kc.nis2 = dec(hb.r_s[0]); kc.nig2 = dec(hb.r_g[0]); kc.nig3 = dec(hb.N0g_exp[0]);
  kc.niIN2 = dec(Nt_IN[0]);
  kc.r_c1 = hb.r_c[0]; kc.r_i1 = hb.r_i[0]; kc.r_r1 = hb.r_r[0]; kc.r_s1 = hb.r_s[0]; kc.r_g1 = hb.r_g[0];
  kc.Nt_i1 = hb.Nt_i[0];
  for (int n = -32; n < 32; ++n) kc.p10[n + 32] = powi10(n);

  // size bins, M:604-670
  hb.Dc[0] = (double)KP_D0C * 1.0; hb.dtc[0] = (double)KP_D0C * 1.0;
  for (int n = 1; n < NBINS; ++n) { hb.Dc[n] = hb.Dc[n - 1] + 1.0E-6; hb.dtc[n] = hb.Dc[n] - hb.Dc[n - 1]; }
  bins(wpd, (double)kc.D0i * 1.0, 5.0 * (double)KP_D0S, hb.Di, hb.dti);
  bins(wpd, (double)KP_D0R * 1.0, 0.005, hb.Dr, hb.dtr);
  bins(wpd, (double)KP_D0S * 1.0, 0.02, hb.Ds, hb.dts);
  bins(wpd, (double)KP_D0G * 1.0, 0.05, hb.Dg, hb.dtg);
  bins(wpd, 1.0, 3000.0, hb.t_Nc, nullptr);
  for (int n = 0; n < NBINS; ++n) hb.t_Nc[n] *= 1.E6;
  kc.nic1 = (int)std::log(hb.t_Nc[NBINS - 1] / hb.t_Nc[0]);               // M:670 (integer assignment truncates)
  kc.Dr1 = hb.Dr[0]; kc.Ds1 = hb.Ds[0];
  kc.lnDr = std::log(hb.Dr[NBINS - 1] / hb.Dr[0]); kc.lnDs = std::log(hb.Ds[NBINS - 1] / hb.Ds[0]);

  kc.n0r_fac = std::pow(kc.crg[2] * kc.org2 * kc.org1, bm_r);
  kc.n0g_fac = std::pow(kc.cgg[2] * kc.ogg2 * kc.ogg1, bm_g);
  kc.lamg_fac = std::pow(kc.cgg[2] * kc.ogg2 * kc.ogg1, kc.obmg);
  for (int n = 0; n < 15; ++n) kc.dcg_fac[n] = std::pow(kc.ccg[2][n] * kc.ocg2[n], kc.obmr);

  // ---- per-axis-node scalars of the table builders ------------------------------------------------
  pp.am_s = am_s;
  pp.lamr.resize(NTB_R * NTB_R1); pp.N0_r.resize(NTB_R * NTB_R1);
  for (int m = 0; m < NTB_R; ++m)                                           // M:3755-3757, M:4126-4128
    for (int k = 0; k < NTB_R1; ++k) {
      const double lam_exp = (double)std::pow(hb.N0r_exp[k] * am_r * kc.crg[0] / hb.r_r[m], kc.ore1);
      const double lamr = lam_exp * (double)std::pow(kc.crg[2] * kc.org2 * kc.org1, kc.obmr);
      pp.lamr[m * NTB_R1 + k] = lamr;
      pp.N0_r[m * NTB_R1 + k] = (double)hb.N0r_exp[k] / ((double)kc.crg[1] * lam_exp) * std::pow(lamr, (double)kc.cre[1]);
    }
  pp.lamg.resize(NTB_G * NTB_G1); pp.N0_g.resize(NTB_G * NTB_G1);
  for (int j = 0; j < NTB_G; ++j)                                           // M:3764-3766
    for (int i = 0; i < NTB_G1; ++i) {
      const double lam_exp = (double)std::pow(hb.N0g_exp[i] * am_g * kc.cgg[0] / hb.r_g[j], kc.oge1);
      const double lamg = lam_exp * (double)std::pow(kc.cgg[2] * kc.ogg2 * kc.ogg1, kc.obmg);
      pp.lamg[j * NTB_G1 + i] = lamg;
      pp.N0_g[j * NTB_G1 + i] = (double)hb.N0g_exp[i] / ((double)kc.cgg[1] * lam_exp) * std::pow(lamg, (double)kc.cge[1]);
    }
  // snow: Field et al. moments at the nine table temperatures, M:3937-3971 (U8 kept as written)
  static const float Tc[NTB_T] = {-0.01f, -5.f, -10.f, -15.f, -20.f, -25.f, -30.f, -35.f, -40.f};
  static const float sa[10] = {5.065339f, -0.062659f, -3.032362f, 0.029469f, -0.000285f, 0.31255f, 0.000204f, 0.003199f, 0.0f, -0.015952f};
  static const float sb[10] = {0.476221f, -0.015896f, 0.165977f, 0.007468f, -0.000141f, 0.060366f, 0.000079f, 0.000594f, 0.0f, -0.003577f};
  auto field = [&](float tc0, float c, float& la, float& b) {
    la = sa[0] + sa[1] * tc0 + sa[2] * c + sa[3] * tc0 * c + sa[4] * tc0 * tc0 + sa[5] * c * c + sa[6] * tc0 * tc0 * c
         + sa[7] * tc0 * c * c + sa[8] * tc0 * tc0 * tc0 + sa[9] * c * c * c;
    b = sb[0] + sb[1] * tc0 + sb[2] * c + sb[3] * tc0 * c + sb[4] * tc0 * tc0 + sb[5] * c * c + sb[6] * tc0 * tc0 * c
        + sb[7] * tc0 * c * c + sb[8] * tc0 * tc0 * tc0 + sb[9] * c * c * c;
  };
  pp.s_Mrat.resize(NTB_T * NTB_S); pp.s_M0.resize(NTB_T * NTB_S); pp.s_slam1.resize(NTB_T * NTB_S); pp.s_slam2.resize(NTB_T * NTB_S);
  for (int j = 0; j < NTB_T; ++j)
    for (int i = 0; i < NTB_S; ++i) {
      const double M2 = (double)(hb.r_s[i] * kc.oams) * 1.0;
      float la, b;
      double second = M2;
      if (bm_s > 2.0f - 1.E-3f && bm_s < 2.0f + 1.E-3f) {
        field(Tc[j], bm_s, la, b);
        second = std::pow(M2 / std::pow((double)10.0f, (double)la), (double)1.f / (double)b);
      }
      field(Tc[j], kc.cse[0], la, b);
      const double M3 = std::pow((double)10.0f, (double)la) * std::pow(second, (double)b);
      const double oM3 = (double)1.f / M3;
      const int ij = j * NTB_S + i;
      pp.s_Mrat[ij] = M2 * (M2 * oM3) * (M2 * oM3) * (M2 * oM3);
      pp.s_M0[ij] = std::pow(M2 * oM3, (double)mu_s);
      pp.s_slam1[ij] = M2 * oM3 * (double)KP_LAM0;
      pp.s_slam2[ij] = M2 * oM3 * (double)KP_LAM1;
    }
  // freezeH2O at m = ntb_IN (U4), M:4118-4122, M:4155-4158
  {
    const float T_adjust = std::fmax(-3.0f, std::fmin(3.0f - std::log10(Nt_IN[NTB_I1 - 1]), 3.0f));
    for (int k = 1; k <= NTB_TC; ++k) {
      const double kk = wpd ? (double)k : (double)(float)k;
      pp.Texp[k - 1] = std::exp(kk - (double)T_adjust * 1.0) - 1.0;
    }
    const int nu_c = std::min(15, (int)std::lround((double)1000.E6f / hb.t_Nc[0]) + 2);
    pp.nu_c_fz = nu_c;
    pp.lamc.resize(NTB_C); pp.N0_c.resize(NTB_C);
    for (int i = 0; i < NTB_C; ++i) {
      const double lamc = std::pow(hb.t_Nc[0] * (double)am_r * (double)kc.ccg[1][nu_c - 1] * (double)kc.ocg1[nu_c - 1]
                                       / (double)hb.r_c[i], (double)kc.obmr);
      pp.lamc[i] = lamc;
      pp.N0_c[i] = hb.t_Nc[0] * (double)kc.ocg1[nu_c - 1] * std::pow(lamc, (double)kc.cce[0][nu_c - 1]);
    }
  }
  // qi_aut_qs node scalars, M:4202-4219
  pp.i_lami.resize(N_IAUS); pp.i_N0.resize(N_IAUS); pp.i_tpi_ide.resize(N_IAUS); pp.i_branch.resize(N_IAUS);
  for (int j = 0; j < NTB_I1; ++j)
    for (int i = 0; i < NTB_I; ++i) {
      const int t = j * NTB_I + i;
      const double lami = (double)std::pow(am_i * kc.cig[1] * kc.oig1 * hb.Nt_i[j] / hb.r_i[i], kc.obmi);
      const double Di_mean = (double)(bm_i + mu_i + 1.f) / lami;
      pp.i_lami[t] = lami;
      pp.i_N0[t] = (double)(hb.Nt_i[j] * kc.oig1) * std::pow(lami, (double)kc.cie[0]);
      if ((float)Di_mean > 5.f * KP_D0S) { pp.i_branch[t] = 0; pp.i_tpi_ide[t] = 0.0; }
      else if ((float)Di_mean < kc.D0i) { pp.i_branch[t] = 1; pp.i_tpi_ide[t] = 1.0; }
      else {
        pp.i_branch[t] = 2;
        pp.i_tpi_ide[t] = (double)gammp(mu_i + 2.0f, (float)(lami * (double)KP_D0S)) * 1.0;
      }
    }
}

}  // namespace hostinit
}  // namespace kidmp
