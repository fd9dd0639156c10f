// kidmp_column.cuh - the Thompson column step on the device, one thread per column.
//
// Replaces the body of `do i = 1, nx` around `call mp_thompson` (I:54-246) and mp_thompson itself
// (M:1156-3688).  Data layout: every field is [nz][ncol] f32 (columns fastest), so the 32 lanes
// of a warp read 128 contiguous bytes per level.
//
// Kernels:
//   k_classify     one bottom-up read of the ten fields decides `no_micro` (M:1396-1521, the early RETURN
//                  at M:1540); clear-sky columns only get back the species <= R1 that the reference
//                  zeroes in the caller's arrays (M:1412-1489) and are done.
//   k_list_scan /  the ballots of the cloudy lanes become a compacted work list in column order.
//   k_list_fill
//   k_column_step  ONE top-down sweep over the cloudy columns does stages S1..S13 of SURVEY.md section 3.2
//                  level by level: every vertical dependency of the scheme runs from the top (graupel N0
//                  running minimum M:1648, `k_0` M:1635, fall-speed carry-down M:3235), so it is carried
//                  along the sweep.  24 values per level are handed to the sedimentation kernel.
//                  (kidmp_units.cuh holds k_unit_step, the same cell code with (32 columns x ONE level) as a warp's
//                  unit of work: the kernel of choice for domains that cannot fill the GPU with a column walk.)
//   k_sediment     sub-stepped upwind sedimentation (M:3365-3578), instant melt/freeze (M:3584-3606),
//                  apply tendencies and final clamps (M:3623-3686), coalesced stores.
//   k_diag_columns the eight domain sums in column order (bitwise reproducible), k_diag_reduce adds the blocks.
#pragma once
#include "kidmp_internal.h"
#include "kidmp_math.cuh"

namespace kidmp {

__constant__ KConst ck;

// The f64 logarithm / exponential / power stay out of line (kidmp_math.cuh): ~130 call sites, and in line they made a
// kernel whose instruction fetch was the limit (profiles/r01; still 6 % slower in line under the lockstep barriers).
// The small helpers below are in line again since the lockstep blocks share their instruction-cache lines: measured
// 2 % faster than out of line (a call costs two branch latencies on a path that is latency-bound).
#ifndef KIDMP_HELPER
#define KIDMP_HELPER __device__ __forceinline__
#endif

#define R1 KP_R1
#define R2 KP_R2
#define EPSF KP_EPS
#define T_0 KP_T_0
#define D0r KP_D0R
#define D0c KP_D0C
#define D0s KP_D0S
#define D0g KP_D0G

// Field et al. (2005) moment relation coefficients, M:305-312
__device__ __constant__ float c_sa[11] = {0, 5.065339f, -0.062659f, -3.032362f, 0.029469f, -0.000285f,
                                          0.31255f, 0.000204f, 0.003199f, 0.0f, -0.015952f};
__device__ __constant__ float c_sb[11] = {0, 0.476221f, -0.015896f, 0.165977f, 0.007468f, -0.000141f,
                                          0.060366f, 0.000079f, 0.000594f, 0.0f, -0.003577f};

// loga_/b_ for moment order c at tc0, M:1590-1599 (term order kept)
__device__ __forceinline__ void field_ab(float tc0, float c, float& loga_, float& b_) {
  const float* sa = c_sa; const float* sb = c_sb;
  loga_ = sa[1] + sa[2] * tc0 + sa[3] * c + sa[4] * tc0 * c + sa[5] * tc0 * tc0 + sa[6] * c * c
          + sa[7] * tc0 * tc0 * c + sa[8] * tc0 * c * c + sa[9] * tc0 * tc0 * tc0 + sa[10] * c * c * c;
  b_ = sb[1] + sb[2] * tc0 + sb[3] * c + sb[4] * tc0 * c + sb[5] * tc0 * tc0 + sb[6] * c * c
       + sb[7] * tc0 * tc0 * c + sb[8] * tc0 * c * c + sb[9] * tc0 * tc0 * tc0 + sb[10] * c * c * c;
}
KIDMP_HELPER float field_moment(float tc0, float c, float smo2) {
  float loga_, b_;
  field_ab(tc0, c, loga_, b_);
  return pow10_f(loga_) * pow_f(smo2, b_);
}

// decade-mantissa table index, M:1762-1774 (f32) / M:1824-1833 (f64 argument).  The reference
// starts its three-candidate search at NINT(log10 x); the candidate that matches does not depend
// on how that logarithm rounds, so a bit-level estimate is enough.
__device__ __forceinline__ int decade_guess(float x) {
  const int b = __float_as_int(x);
  const float l2 = (float)((b >> 23) - 127) + __int_as_float((b & 0x007fffff) | 0x3f800000) - 1.0f;
  return __float2int_rn(l2 * 0.30103f);
}
// When x / 10.**n0 is well inside [1, 10) no other candidate can match (x / 10.**(n0-1) is ten times larger, x / 10.**(n0+1)
// ten times smaller, and a division is off by half an ulp at most), so the search of the reference is only walked,
// in its own order, when the quotient is within 1e-5 of a decade boundary or the guess is a decade off.
KIDMP_HELPER int decade_idx_f(float x, int n2, int ntb) {
  const int n0 = decade_guess(x);
  const float q0 = x / ck.p10[n0 + 32];
  int n = n0;
  float qn = q0;
  if (!(q0 > 1.00001f && q0 < 9.9999f)) {
    n = n0 + 1;
#pragma unroll
    for (int nn = -1; nn <= 1; ++nn) {
      const float q = x / ck.p10[n0 + nn + 32];
      if (q >= 1.0f && q < 10.0f) { n = n0 + nn; break; }
    }
    qn = x / ck.p10[n + 32];
  }
  const int idx = (int)qn + 9 * (n - n2);
  return max(1, min(idx, ntb));
}
KIDMP_HELPER int decade_idx_d(double x, int n2, int ntb) {
  const int n0 = decade_guess((float)x);
  const double q0 = x / (double)ck.p10[n0 + 32];
  int n = n0;
  double qn = q0;
  if (!(q0 > 1.00001 && q0 < 9.9999)) {
    n = n0 + 1;
#pragma unroll
    for (int nn = -1; nn <= 1; ++nn) {
      const double q = x / (double)ck.p10[n0 + nn + 32];
      if (q >= 1.0 && q < 10.0) { n = n0 + nn; break; }
    }
    qn = x / (double)ck.p10[n + 32];
  }
  const int idx = (int)qn + 9 * (n - n2);
  return max(1, min(idx, ntb));
}

// rain number from mass at a clamped median volume diameter, M:1452-1454
KIDMP_HELPER float nr_from_mvd(float rr, float mvd_r) {
  const double lamr = (double)((3.0f + 0.0f + 0.672f) / mvd_r);
  return (float)((double)(ck.crg[1] * ck.org3 * rr) * cube_d(lamr) / (double)ck.am_r);
}
// lamr = (am_r*crg(3)*org2*nr/rr)**obmr, M:1457
KIDMP_HELPER double rain_lam(float nr, float rr) {
  return (double)pow_f(ck.am_r * ck.crg[2] * ck.org2 * nr / rr, ck.obmr);
}
KIDMP_HELPER double ice_lam(float ni, float ri) {   // M:1429
  return (double)pow_f(ck.am_i * ck.cig[1] * ck.oig1 * ni / ri, ck.obmi);
}

// graupel intercept at one level given the running minimum from above, M:1639-1653.
// xslw1 = 0.01 and MAX(5.E-5, rg) = 5.E-5 (no supercooled rain above k_0, graupel content at most 5.E-5) make
// N0_exp the per-run constant `n0_lo` (evaluated once per thread by graupel_n0_lo with the same expressions);
// lam_exp, lamg, ilamg and N0_g (M:1649-1653) are only read where rg > R1 (M:1862, M:1921, M:2168, M:2254, M:3318).
__device__ __forceinline__ double graupel_n0_exp(float xslw1, float rg) {
  const float ygra1 = 4.31f + log10_f(fmaxf(5.E-5f, rg));
  const float zans1 = 3.1f + (100.f / (300.f * xslw1 * ygra1 / (10.f / xslw1 + 1.f + 0.25f * ygra1) + 30.f + 10.f * ygra1));
  const double N0_exp = (double)pow10_f(zans1);
  return fmax((double)KP_GONV_MIN, fmin(N0_exp, (double)KP_GONV_MAX));
}
KIDMP_HELPER double graupel_n0_lo() { return graupel_n0_exp(0.01f, R1); }
__device__ double g_n0_lo;                                 // graupel_n0_lo(), evaluated once at init by k_n0_lo
__global__ void k_n0_lo() { g_n0_lo = graupel_n0_lo(); }
KIDMP_HELPER void graupel_n0(bool above_k0, bool L_qr, bool L_qg, float mvd_r, float rg, double n0_lo, double& N0_min,
                             double& ilamg, double& N0_g) {
  const bool slw = above_k0 && L_qr && mvd_r > 100.E-6f;
  double N0_exp = n0_lo;
  if (slw || rg > 5.E-5f) {
    float xslw1 = 0.01f;
    if (slw) xslw1 = 4.01f + log10_f(mvd_r);
    N0_exp = graupel_n0_exp(xslw1, rg);
  }
  N0_min = fmin(N0_exp, N0_min);
  N0_exp = N0_min;
  if (L_qg) {
    const double lam_exp = sqrt(sqrt(N0_exp * (double)ck.am_g * (double)ck.cgg[0] / (double)rg));   // **oge1, oge1 = 1/4
    const double lamg = lam_exp * (double)ck.lamg_fac;
    ilamg = (double)1.f / lamg;
    N0_g = N0_exp / ((double)ck.cgg[1] * lam_exp) * lamg;                                         // lamg**cge(2), cge(2) = 1
  }
}

enum { F_QV = 0, F_QC, F_QI, F_QR, F_QS, F_QG, F_NI, F_NR, F_T };

// ---- K0: classification (pass 0).  One thread per column reads the ten fields once, decides
// `no_micro` (M:1396-1521; the early RETURN at M:1540), writes back the species <= R1 that the
// reference zeroes in the caller's arrays before returning (M:1412-1489, U9), marks clear-sky columns
// in colint[0] and appends every 32-column group that holds a cloudy column to the work list of the
// physics kernel.  Light on registers: many warps per SM keep the HBM pipe full for the ~70 % of
// columns that need nothing else.
__global__ void __launch_bounds__(128, 8) k_classify(StepArgs a) {
  const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = col < a.ncol;
  const int nz = a.nz;
  const long ncol = a.ncol;
  bool active = false;
  if (in_range) {
    const float* __restrict__ Gp = a.p + col;
    float* Gqv = a.f[F_QV] + col; float* Gqc = a.f[F_QC] + col; float* Gqi = a.f[F_QI] + col;
    float* Gqr = a.f[F_QR] + col; float* Gqs = a.f[F_QS] + col; float* Gqg = a.f[F_QG] + col;
    float* Gni = a.f[F_NI] + col; float* Gnr = a.f[F_NR] + col; float* Gt = a.f[F_T] + col;
    bool no_micro = true;
#pragma unroll 4
    for (int k = 0; k < nz; ++k) {
      const long o = (long)k * ncol;
      const float qc = Gqc[o], qi = Gqi[o], qr = Gqr[o], qs = Gqs[o], qg = Gqg[o];
      const float ni = Gni[o], nr = Gnr[o];
      const float t = Gt[o], pr = Gp[o], qv = fmaxf(1.E-10f, Gqv[o]);
      if (qc > R1 || qi > R1 || qr > R1 || qs > R1 || qg > R1) no_micro = false;
      if (!(qc > R1) && qc != 0.0f) Gqc[o] = 0.0f;
      if (!(qi > R1) && (qi != 0.0f || ni != 0.0f)) { Gqi[o] = 0.0f; Gni[o] = 0.0f; }
      if (!(qr > R1) && (qr != 0.0f || nr != 0.0f)) { Gqr[o] = 0.0f; Gnr[o] = 0.0f; }
      if (!(qs > R1) && qs != 0.0f) Gqs[o] = 0.0f;
      if (!(qg > R1) && qg != 0.0f) Gqg[o] = 0.0f;
      const float tempc = t - 273.15f;
      const float qvsi = (tempc <= 0.0f) ? rsif(pr, t) : rslf(pr, t);
      float ssati = qv / qvsi - 1.f;
      if (fabsf(ssati) < EPSF) ssati = 0.0f;
      if (ssati > 0.0f) no_micro = false;
    }
    active = !no_micro;
    if (!active) {                                 // clear-sky column: nothing left to do
      a.colint[col] = -1;
      a.ppt[col] = 0.f; a.ppt[ncol + col] = 0.f; a.ppt[2 * ncol + col] = 0.f; a.ppt[3 * ncol + col] = 0.f;   // I:55-58
    }
  }
  // ballot of the cloudy lanes of this 32-column group; k_list_scan / k_list_fill turn the ballots into the
  // compacted work list IN COLUMN ORDER (neighbouring lanes of the physics kernel are neighbouring columns: coalesced
  // accesses, similar branches, and a list that is identical from run to run)
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if ((threadIdx.x & 31) == 0 && in_range) a.work_mask[col >> 5] = mask;
}

// exclusive prefix sum of the per-group cloudy-column counts (one block; ngroups is at most a few 10^5)
__global__ void __launch_bounds__(1024) k_list_scan(const unsigned* __restrict__ mask, int ngroups, int* __restrict__ offset,
                                                    int* __restrict__ count) {
  __shared__ int s_sum[1024];
  const int per = (ngroups + 1023) / 1024;
  const int g0 = threadIdx.x * per, g1 = min(g0 + per, ngroups);
  int sum = 0;
  for (int g = g0; g < g1; ++g) sum += __popc(mask[g]);
  s_sum[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {                     // Hillis-Steele inclusive scan
    const int v = (int)threadIdx.x >= d ? s_sum[threadIdx.x - d] : 0;
    __syncthreads();
    s_sum[threadIdx.x] += v;
    __syncthreads();
  }
  int run = s_sum[threadIdx.x] - sum;
  for (int g = g0; g < g1; ++g) { offset[g] = run; run += __popc(mask[g]); }
  if (threadIdx.x == 1023) *count = s_sum[1023];
}

__global__ void __launch_bounds__(256) k_list_fill(const unsigned* __restrict__ mask, const int* __restrict__ offset, int ngroups,
                                                   int* __restrict__ list) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= ngroups) return;
  const unsigned m = mask[g];
  if ((m >> lane) & 1u) list[offset[g] + __popc(m & ((1u << lane) - 1u))] = g * 32 + lane;
}

// ---- the last (or only) sedimentation sub-step of the four species at one level (M:3365-3578), the instant melting /
// freezing of cloud ice and cloud water (S15, M:3584-3606), the tendencies applied with the final clamps (S16,
// M:3623-3686), the nine output stores and the water paths of the new state.  Walked top-down: the fluxes of the level
// above come in `c`.  Shared by k_sediment (after the extra sub-steps) and by the fused physics kernel.
struct SedParams {            // per column, fixed over the sweep
  float DT, odt, on_r, on_i, on_s, on_g, Nt_c;
  int top_r, top_i, top_s, top_g;   // ksed1(1..4): top sedimenting level of rain, ice, snow, graupel (M:3208)
  bool sedi, iiwarm;
};
struct SedCarry {             // carried down the column
  float sr_up, snr_up, si_up, sni_up, ss_up, sg_up;     // fluxes leaving the level above
  float ppt_r, ppt_i, ppt_s, ppt_g;
  double lwp, iwp;
};
struct HandOff {              // what S1..S13 leave for a level (SC_* order)
  float tt, qvt, qct, qit, qrt, qst, qgt, nit, nrt, nct, rr, nr, ri, ni, rs, rg, v_r, v_nr, v_i, v_ni, v_s, v_g, rho, s15;
};
__device__ __forceinline__ void finish_level(const StepArgs& a, const SedParams& p, SedCarry& c, const HandOff& h, int k, int nz,
                                             long g, float dzk, float t1d, float qv1d, float qc1d, float qi1d, float qr1d,
                                             float qs1d, float qg1d, float ni1d, float nr1d, float pres) {
  float tt = h.tt, qvt = h.qvt, qct = h.qct, qit = h.qit, qrt = h.qrt, qst = h.qst, qgt = h.qgt, nit = h.nit, nrt = h.nrt, nct = h.nct;
  const float rho = h.rho, s15 = h.s15, rr = h.rr, nr = h.nr, ri = h.ri, ni = h.ni, rs = h.rs, rg = h.rg;
  const float odzq = 1.f / dzk, orho = 1.f / rho;
  const float sr = h.v_r * rr, snr = h.v_nr * nr;
  const float si = p.sedi ? h.v_i * ri : 0.f, sni = p.sedi ? h.v_ni * ni : 0.f;
  const float ssn = p.sedi ? h.v_s * rs : 0.f, sg = p.sedi ? h.v_g * rg : 0.f;
  float rr_n = rr, ri_n = ri, rs_n = rs, rg_n = rg;
  if (k == nz - 1) {
    qrt = qrt - sr * odzq * p.on_r * orho;   nrt = nrt - snr * odzq * p.on_r * orho;
    rr_n = fmaxf(KP_R1, rr - sr * odzq * p.DT * p.on_r);
    qit = qit - si * odzq * p.on_i * orho;   nit = nit - sni * odzq * p.on_i * orho;
    ri_n = fmaxf(KP_R1, ri - si * odzq * p.DT * p.on_i);
    qst = qst - ssn * odzq * p.on_s * orho;  rs_n = fmaxf(KP_R1, rs - ssn * odzq * p.DT * p.on_s);
    qgt = qgt - sg * odzq * p.on_g * orho;   rg_n = fmaxf(KP_R1, rg - sg * odzq * p.DT * p.on_g);
  } else {
    if (k + 1 <= p.top_r) {
      qrt = qrt + (c.sr_up - sr) * odzq * p.on_r * orho;   nrt = nrt + (c.snr_up - snr) * odzq * p.on_r * orho;
      rr_n = fmaxf(KP_R1, rr + (c.sr_up - sr) * odzq * p.DT * p.on_r);
    }
    if (k + 1 <= p.top_i) {
      qit = qit + (c.si_up - si) * odzq * p.on_i * orho;   nit = nit + (c.sni_up - sni) * odzq * p.on_i * orho;
      ri_n = fmaxf(KP_R1, ri + (c.si_up - si) * odzq * p.DT * p.on_i);
    }
    if (k + 1 <= p.top_s) {
      qst = qst + (c.ss_up - ssn) * odzq * p.on_s * orho;  rs_n = fmaxf(KP_R1, rs + (c.ss_up - ssn) * odzq * p.DT * p.on_s);
    }
    if (k + 1 <= p.top_g) {
      qgt = qgt + (c.sg_up - sg) * odzq * p.on_g * orho;   rg_n = fmaxf(KP_R1, rg + (c.sg_up - sg) * odzq * p.DT * p.on_g);
    }
  }
  c.sr_up = sr; c.snr_up = snr; c.si_up = si; c.sni_up = sni; c.ss_up = ssn; c.sg_up = sg;
  if (k == 0) {                                         // surface precipitation of the last sub-step, M:3391-3392
    if (rr_n > KP_R1 * 10.f) c.ppt_r = c.ppt_r + sr * p.DT * p.on_r;
    if (ri_n > KP_R1 * 10.f) c.ppt_i = c.ppt_i + si * p.DT * p.on_i;
    if (rs_n > KP_R1 * 10.f) c.ppt_s = c.ppt_s + ssn * p.DT * p.on_s;
    if (rg_n > KP_R1 * 10.f) c.ppt_g = c.ppt_g + sg * p.DT * p.on_g;
  }

  // ---- S15 + S16 for this level ----------------------------------------------------------------
  float nc1d = p.Nt_c / (0.622f * pres / (KP_R * t1d * (qv1d + 0.622f)));   // U1
  if (!(qc1d > KP_R1)) { qc1d = 0.f; nc1d = 0.f; }
  if (!(qi1d > KP_R1)) { qi1d = 0.f; ni1d = 0.f; }
  if (!(qr1d > KP_R1)) { qr1d = 0.f; nr1d = 0.f; }
  if (!(qs1d > KP_R1)) qs1d = 0.f;
  if (!(qg1d > KP_R1)) qg1d = 0.f;
  if (!p.iiwarm) {
    const float xri = fmaxf(0.0f, qi1d + qit * p.DT);
    if ((s15 > 0.f) && (xri > 0.0f)) {                  // temp > T_0
      qct = qct + xri * p.odt;
      nct = nct + ni1d * p.odt;
      qit = qit - xri * p.odt;
      nit = -ni1d * p.odt;
      tt = tt - s15 * xri * p.odt * 1.0f;
    }
    const float xrc = fmaxf(0.0f, qc1d + qct * p.DT);
    if ((s15 < 0.f) && (xrc > 0.0f)) {                  // temp < HGFR
      const float xnc = nc1d + nct * p.DT;
      qit = qit + xrc * p.odt;
      nit = nit + xnc * p.odt;
      qct = qct - xrc * p.odt;
      nct = nct - xnc * p.odt;
      tt = tt + (-s15) * xrc * p.odt * 1.0f;
    }
  }
  t1d = t1d + tt * p.DT;
  qv1d = fmaxf(1.E-10f, qv1d + qvt * p.DT);
  qc1d = qc1d + qct * p.DT;
  if (qc1d <= KP_R1) qc1d = 0.0f;              // nc1d is not returned to the host (I:143-152, I:198-245)
  qi1d = qi1d + qit * p.DT;
  ni1d = fmaxf(KP_R2 / rho, ni1d + nit * p.DT);
  if (qi1d <= KP_R1) {
    qi1d = 0.0f; ni1d = 0.0f;
  } else {
    double lami = ice_lam(ni1d, qi1d);
    const double ilami = (double)1.f / lami;
    const float xDi = (float)((double)(3.f + 0.f + 1.f) * ilami);
    if (xDi < 5.E-6f) lami = (double)(ck.cie[1] / 5.E-6f);
    else if (xDi > 300.E-6f) lami = (double)(ck.cie[1] / 300.E-6f);
    ni1d = (float)fmin((double)(ck.cig[0] * ck.oig2 * qi1d / ck.am_i) * cube_d(lami), 499.E3 / (double)rho);
  }
  qr1d = qr1d + qrt * p.DT;
  nr1d = fmaxf(KP_R2 / rho, nr1d + nrt * p.DT);
  if (qr1d <= KP_R1) {
    qr1d = 0.0f; nr1d = 0.0f;
  } else {
    const double lamr = rain_lam(nr1d, qr1d);
    float mvd_r = (float)((double)(3.0f + 0.f + 0.672f) / lamr);
    if (mvd_r > 2.5E-3f) mvd_r = 2.5E-3f;
    else if (mvd_r < KP_D0R * 0.75f) mvd_r = KP_D0R * 0.75f;
    nr1d = nr_from_mvd(qr1d, mvd_r);
  }
  qs1d = qs1d + qst * p.DT;
  if (qs1d <= KP_R1) qs1d = 0.0f;
  qg1d = qg1d + qgt * p.DT;
  if (qg1d <= KP_R1) qg1d = 0.0f;
  a.f[F_T][g] = t1d; a.f[F_QV][g] = qv1d; a.f[F_QC][g] = qc1d; a.f[F_QI][g] = qi1d; a.f[F_QR][g] = qr1d; a.f[F_QS][g] = qs1d; a.f[F_QG][g] = qg1d;
  a.f[F_NI][g] = ni1d; a.f[F_NR][g] = nr1d;
  // domain diagnostics: liquid / ice water paths of the new state
  const float rho_new = 0.622f * pres / (KP_R * t1d * (qv1d + 0.622f));
  c.lwp += (double)((qc1d + qr1d) * rho_new * dzk);
  c.iwp += (double)((qi1d + qs1d + qg1d) * rho_new * dzk);
}

// ---- K1: column physics, S1..S13, on the cloudy 32-column groups of the work list.  A block is
// WARPS warps = WARPS groups; its warps meet at a named barrier at every level of the top-down sweep,
// so they run the same ~120 KB of straight-line code at the same time and share its instruction-cache
// lines (profiles/r01: with independent warps the GPC instruction cache sat at 98 % of its request
// peak and `no_instruction` was 8 of 13 stall cycles per issue).
// FUSE: the single-sub-step sedimentation, S15 and S16 of a level follow its S1..S13 at once (finish_level) and the
// new state is stored in place, so the 24-value hand-off is neither written nor read and k_sediment does not run.
// That is only right for columns whose four sub-step counts end up <= 1, which is known at the bottom of the sweep:
// the inputs of every level are therefore parked in the (otherwise unused) hand-off buffer first, and a column that
// turns out to need sub-steps is put on the redo list - k_restore brings its inputs back and the split kernels
// (FUSE = false, then k_sediment) do it again.
template <int WARPS, int MINB, int BARS, bool RATES, bool FUSE>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_column_step(StepArgs a) {
  const int count = *a.work_count;                     // cloudy columns, compacted: every warp but the last is full
  // The warps of the list are dealt evenly to a whole number of waves of blocks (one wave = MINB blocks on each SM):
  // a block runs for milliseconds, so a last wave that fills only part of the SMs would leave the others idle for
  // that long.  Blocks get WARPS or fewer warps; surplus blocks and warps leave at once.
  const int total_warps = (count + 31) >> 5;
  int nblocks = (total_warps + WARPS - 1) / WARPS;
  const int wave = a.nsm * MINB;
  nblocks = min((int)gridDim.x, (nblocks + wave - 1) / wave * wave);
  if ((int)blockIdx.x >= nblocks) return;
  const int w0 = (int)((long)blockIdx.x * total_warps / nblocks), w1 = (int)((long)(blockIdx.x + 1) * total_warps / nblocks);
  const int mywarp = threadIdx.x >> 5;
  if (mywarp >= w1 - w0) return;
  const int lock_threads = (w1 - w0) * 32;
  const int wfirst = (w0 + mywarp) * 32;
  const int slot = wfirst + (threadIdx.x & 31);
  const bool active = slot < count;
  // lanes past the end of the list shadow the warp's first column: same branches, nothing stored, and the warp
  // stays convergent at the stage barriers
  const long col = (long)a.work_list[active ? slot : wfirst];
  const int nz = a.nz;
  const long ncol = a.ncol;
  const float DT = a.dt;
  const float odt = 1.f / DT, odts = 1.f / DT;
  const float Nt_c = ck.Nt_c;
  const bool iiwarm = ck.iiwarm != 0;
  constexpr bool LOCKSTEP = WARPS > 1;
  // stage barrier of the lockstep block: the warps stay within one instruction-cache window of each other
#define LOCKBAR(i) do { if (LOCKSTEP && ((BARS >> (i)) & 1)) { __syncwarp(); asm volatile("bar.sync 1, %0;" ::"r"(lock_threads) : "memory"); } } while (0)
  {
    const long colc = col;
    const float* __restrict__ Gp = a.p + colc;
    float* Gqv = a.f[F_QV] + colc; float* Gqc = a.f[F_QC] + colc; float* Gqi = a.f[F_QI] + colc;
    float* Gqr = a.f[F_QR] + colc; float* Gqs = a.f[F_QS] + colc; float* Gqg = a.f[F_QG] + colc;
    float* Gni = a.f[F_NI] + colc; float* Gnr = a.f[F_NR] + colc; float* Gt = a.f[F_T] + colc;
    {
      // carried from the level above
      // They are touched once per level and live for the whole sweep: kept in shared memory (72 bytes per
      // thread) instead of 20 registers that the register allocator would have to spill around the rates.
      // 84 bytes per thread: 768 threads fit the 64 KB shared-memory carve-out and leave 192 KB of the SM to L1, which
      // the register spills and table gathers of the cell code live on
      extern __shared__ double smem_carry[];
      constexpr int NT = WARPS * 32;
      float* const s_in = reinterpret_cast<float*>(smem_carry);         // [9][NT] this level's inputs, parked over S3..S7
      float* const s_n0 = s_in + 9 * NT;                                // [2][NT] graupel N0 running minima (f32 numbers, M:1646-1648)
      float* const s_f = s_n0 + 2 * NT;                                 // [6][NT] fall speeds of the level above
      unsigned short* const s_i = reinterpret_cast<unsigned short*>(s_f + 6 * NT);   // [8][NT] sub-step counts (capped), top sedimenting levels
      double* const s_wp = reinterpret_cast<double*>(s_i + 8 * NT);     // FUSE: [2][NT] liquid / ice water path so far
      float* const s_flux = reinterpret_cast<float*>(s_wp + 2 * NT);    // FUSE: [6][NT] sedimentation fluxes of the level above
      const int tid = threadIdx.x;
#define N0_min_a s_n0[tid]
#define N0_min_b s_n0[NT + tid]
#define vtr_up s_f[tid]
#define vtnr_up s_f[NT + tid]
#define vti_up s_f[2 * NT + tid]
#define vtni_up s_f[3 * NT + tid]
#define vts_up s_f[4 * NT + tid]
#define vtg_up s_f[5 * NT + tid]
#define nstep_r s_i[tid]
#define nstep_i s_i[NT + tid]
#define nstep_s s_i[2 * NT + tid]
#define nstep_g s_i[3 * NT + tid]
#define ksed_r s_i[4 * NT + tid]
#define ksed_i s_i[5 * NT + tid]
#define ksed_s s_i[6 * NT + tid]
#define ksed_g s_i[7 * NT + tid]
      N0_min_a = KP_GONV_MAX; N0_min_b = KP_GONV_MAX;
      bool warm_above_a = false, warm_above_b = false;     // any level >= k with temp >= 270.65 (k_0, M:1635)
      vtr_up = 0.f; vtnr_up = 0.f; vti_up = 0.f; vtni_up = 0.f; vts_up = 0.f; vtg_up = 0.f;
      nstep_r = 0; nstep_i = 0; nstep_s = 0; nstep_g = 0;
      ksed_r = 1; ksed_i = 1; ksed_s = 1; ksed_g = 1;      // 1-based like the reference
      if (FUSE) {
        s_wp[tid] = 0.0; s_wp[NT + tid] = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) s_flux[q * NT + tid] = 0.f;
      }
      // S14..S16 of one level right after its S1..S13 (FUSE)
      auto finish = [&](const HandOff& h, int k, long o, float dzq, float t1d, float qv1d, float qc1d, float qi1d, float qr1d,
                        float qs1d, float qg1d, float ni1d, float nr1d, float pres) {
        SedParams sp;
        sp.DT = DT; sp.odt = odt; sp.on_r = 1.f; sp.on_i = 1.f; sp.on_s = 1.f; sp.on_g = 1.f; sp.Nt_c = Nt_c;
        sp.top_r = ksed_r; sp.top_i = ksed_i; sp.top_s = ksed_s; sp.top_g = ksed_g;   // so far = final for this level
        sp.sedi = ck.l_sediment != 0; sp.iiwarm = iiwarm;
        SedCarry c;
        c.sr_up = s_flux[tid]; c.snr_up = s_flux[NT + tid]; c.si_up = s_flux[2 * NT + tid]; c.sni_up = s_flux[3 * NT + tid];
        c.ss_up = s_flux[4 * NT + tid]; c.sg_up = s_flux[5 * NT + tid];
        c.ppt_r = 0.f; c.ppt_i = 0.f; c.ppt_s = 0.f; c.ppt_g = 0.f;
        c.lwp = s_wp[tid]; c.iwp = s_wp[NT + tid];
        finish_level(a, sp, c, h, k, nz, o + col, dzq, t1d, qv1d, qc1d, qi1d, qr1d, qs1d, qg1d, ni1d, nr1d, pres);
        s_flux[tid] = c.sr_up; s_flux[NT + tid] = c.snr_up; s_flux[2 * NT + tid] = c.si_up; s_flux[3 * NT + tid] = c.sni_up;
        s_flux[4 * NT + tid] = c.ss_up; s_flux[5 * NT + tid] = c.sg_up;
        s_wp[tid] = c.lwp; s_wp[NT + tid] = c.iwp;
        if (k == 0) {     // ppt is overwritten with this step's amounts: rain, ice, snow, graupel (I:55-58, I:162-177)
          a.ppt[col] = c.ppt_r; a.ppt[ncol + col] = c.ppt_i; a.ppt[2 * ncol + col] = c.ppt_s; a.ppt[3 * ncol + col] = c.ppt_g;
          a.coldiag[col] = c.lwp; a.coldiag[ncol + col] = c.iwp;
        }
      };

      // graupel intercept of a level without rain and graupel (xslw1 = 0.01, rg = R1 in M:1639-1646): the only
      // thing such a level contributes to the running minimum of M:1648.  A per-run constant: k_n0_lo evaluates it once
      // at init with the routines used here, and the sweep reads it where needed instead of holding it in registers.
#define n0_empty g_n0_lo

      // ================= pass 1: top-down, S1..S13 per level ====================================
#pragma unroll 1
      for (int k = nz - 1; k >= 0; --k) {
        LOCKBAR(0);
        {
        const long o = (long)k * ncol;
        const float t1d = Gt[o], qv1d = Gqv[o], pres = Gp[o];
        float qc1d = Gqc[o], qi1d = Gqi[o], qr1d = Gqr[o], qs1d = Gqs[o], qg1d = Gqg[o];
        float ni1d = Gni[o], nr1d = Gnr[o];
        const float dzq = a.dz_col ? a.dz_col[o + col] : a.dz[k];
        if (FUSE && active) {                              // the inputs of the level, in case the column must be redone
          float* sc = a.scratch + o + col;
          const long ss = (long)nz * ncol;
          sc[0] = qv1d; sc[ss] = qc1d; sc[2 * ss] = qi1d; sc[3 * ss] = qr1d; sc[4 * ss] = qs1d; sc[5 * ss] = qg1d;
          sc[6 * ss] = ni1d; sc[7 * ss] = nr1d; sc[8 * ss] = t1d;
        }

        // ---- empty level: no hydrometeor in any of the warp's columns and none at or above ice saturation.
        // Every process rate is then exactly zero (each is gated by a species flag or by ssati / ssatw,
        // M:1676-2286, M:2780, M:2880) and the state at tau+1 equals the input, so only the vertical carries
        // (graupel N0 minimum, k_0, fall speeds from above, substep counts) and the hand-off need doing.
        if (!__any_sync(0xffffffffu, qc1d > R1 || qi1d > R1 || qr1d > R1 || qs1d > R1 || qg1d > R1)) {
          const float temp = t1d;
          const float qv = fmaxf(1.E-10f, qv1d);
          const float rho = 0.622f * pres / (KP_R * temp * (qv + 0.622f));
          const float tempc = temp - 273.15f;
          const float qvs = rslf(pres, temp);
          const float qvsi = (tempc <= 0.0f) ? rsif(pres, temp) : qvs;
          float ssatw = qv / qvs - 1.f;
          float ssati = qv / qvsi - 1.f;
          if (fabsf(ssatw) < EPSF) ssatw = 0.0f;
          if (fabsf(ssati) < EPSF) ssati = 0.0f;
          if (__all_sync(0xffffffffu, !(ssati > 0.0f) && !(ssatw > EPSF))) {
            if (!iiwarm) {
              if (temp >= 270.65f) { warm_above_a = true; warm_above_b = true; }
              N0_min_a = (float)fmin(n0_empty, (double)N0_min_a);
              N0_min_b = (float)fmin(n0_empty, (double)N0_min_b);
            }
            const float v_r = vtr_up, v_nr = vtnr_up, v_i = vti_up, v_ni = vtni_up, v_s = vts_up, v_g = vtg_up;
            if (fmaxf(v_r, v_nr) > 1.E-3f) {
              ksed_r = (unsigned short)max((int)ksed_r, k + 1);
              const float delta_tp = dzq / (fmaxf(v_r, v_nr));
              nstep_r = (unsigned short)min(max((int)nstep_r, (int)(DT / delta_tp + 1.f)), KP_NSTEP_MAX);
            }
            if (!iiwarm) {
              if (v_i > 1.E-3f) { ksed_i = (unsigned short)max((int)ksed_i, k + 1); const float d = dzq / v_i; nstep_i = (unsigned short)min(max((int)nstep_i, (int)(DT / d + 1.f)), KP_NSTEP_MAX); }
              if (v_s > 1.E-3f) { ksed_s = (unsigned short)max((int)ksed_s, k + 1); const float d = dzq / v_s; nstep_s = (unsigned short)min(max((int)nstep_s, (int)(DT / d + 1.f)), KP_NSTEP_MAX); }
              if (v_g > 1.E-3f) { ksed_g = (unsigned short)max((int)ksed_g, k + 1); const float d = dzq / v_g; nstep_g = (unsigned short)min(max((int)nstep_g, (int)(DT / d + 1.f)), KP_NSTEP_MAX); }
            }
            if (FUSE && active) {
              const float ocp = 1.f / (KP_CP * (1.f + 0.887f * qv));
              const float lvap = KP_LVAP0 + (2106.0f - 4218.0f) * tempc;
              HandOff h;
              h.tt = 0.f; h.qvt = 0.f; h.qct = 0.f; h.qit = 0.f; h.qrt = 0.f; h.qst = 0.f; h.qgt = 0.f; h.nit = 0.f; h.nrt = 0.f; h.nct = 0.f;
              h.rr = R1; h.nr = R2; h.ri = R1; h.ni = R2; h.rs = R1; h.rg = R1;
              h.v_r = v_r; h.v_nr = v_nr; h.v_i = v_i; h.v_ni = v_ni; h.v_s = v_s; h.v_g = v_g; h.rho = rho;
              h.s15 = 0.0f;
              if (temp > T_0) h.s15 = ck.lfus * ocp;
              else if (temp < KP_HGFR) h.s15 = -((KP_LSUB - lvap) * ocp);
              finish(h, k, o, dzq, t1d, qv1d, qc1d, qi1d, qr1d, qs1d, qg1d, ni1d, nr1d, pres);
            } else if (active) {
              const float ocp = 1.f / (KP_CP * (1.f + 0.887f * qv));
              const float lvap = KP_LVAP0 + (2106.0f - 4218.0f) * tempc;
              float s15 = 0.0f;
              if (temp > T_0) s15 = ck.lfus * ocp;
              else if (temp < KP_HGFR) s15 = -((KP_LSUB - lvap) * ocp);
              float* sc = a.scratch + o + col;
              const long ss = (long)nz * ncol;
#pragma unroll
              for (int q = SC_TTEN; q <= SC_NCTEN; ++q) sc[q * ss] = 0.0f;
              sc[SC_RR * ss] = R1; sc[SC_NR * ss] = R2; sc[SC_RI * ss] = R1; sc[SC_NI * ss] = R2; sc[SC_RS * ss] = R1; sc[SC_RG * ss] = R1;
              sc[SC_VTR * ss] = v_r; sc[SC_VTNR * ss] = v_nr; sc[SC_VTI * ss] = v_i; sc[SC_VTNI * ss] = v_ni;
              sc[SC_VTS * ss] = v_s; sc[SC_VTG * ss] = v_g; sc[SC_RHO * ss] = rho; sc[SC_S15 * ss] = s15;
              if (RATES && a.rates) {
                float* rp = a.rates + o + col;
                for (int q = 0; q < KIDMP_NRATES; ++q) rp[q * ss] = 0.0f;
              }
            }
            LOCKBAR(1); LOCKBAR(2); LOCKBAR(3); LOCKBAR(4); LOCKBAR(5);   // keep the block's barrier count in step
            continue;
          }
        }

        // rates, M:1184-1211 (zeroed M:1282-1363)
        double prw_vcd = 0., pnc_wcd = 0., pnc_wau = 0., pnc_rcw = 0., pnc_scw = 0., pnc_gcw = 0.;
        double prv_rev = 0., prr_wau = 0., prr_rcw = 0., prr_rcs = 0., prr_rcg = 0., prr_sml = 0., prr_gml = 0., prr_rci = 0.;
        double pnr_wau = 0., pnr_rcs = 0., pnr_rcg = 0., pnr_rci = 0., pnr_sml = 0., pnr_gml = 0., pnr_rev = 0., pnr_rcr = 0., pnr_rfz = 0.;
        double pri_inu = 0., pni_inu = 0., pri_ihm = 0., pni_ihm = 0., pri_wfz = 0., pni_wfz = 0., pri_rfz = 0., pni_rfz = 0.;
        double pri_ide = 0., pni_ide = 0., pri_rci = 0., pni_rci = 0., pni_sci = 0., pni_iau = 0.;
        double prs_iau = 0., prs_sci = 0., prs_rcs = 0., prs_scw = 0., prs_sde = 0., prs_ihm = 0., prs_ide = 0.;
        double prg_scw = 0., prg_rfz = 0., prg_gde = 0., prg_gcw = 0., prg_rci = 0., prg_rcs = 0., prg_rcg = 0., prg_ihm = 0.;
        float smo0 = 0.f, smo1 = 0.f, smob = 0.f, smoc = 0.f, smoe = 0.f, smof = 0.f;
        bool have_smoe = false;
        float mvd_r = 0.f, mvd_c = 0.f, vts_boost = 0.f;
        double ilamg = 0., N0_g = 0., ilamr, N0_r, lamr, lamc = 0., lami, ilami;
        int nu_c = 0;
        float xDc = 0.f;
        // number tendencies are summed as their terms appear (M:2417, M:2453, M:2503 add them up later): the
        // 22 individual number rates need not stay in registers until S8
        double nc_acc = 0., ni_acc = 0., nr_acc = 0.;
        // lamr / lami hold rain_lam(nr, rr) / ice_lam(ni, ri) of the current nr, rr / ni, ri unless the number was
        // re-diagnosed after they were evaluated: the reference evaluates the same power again at M:1661, M:2118,
        // M:2750 and M:3227 from unchanged arguments, which is the same number
        bool lamr_stale = false, lami_stale = false;

        // ---- S1, M:1387-1493 -------------------------------------------------------------------
        float temp = t1d;
        float qv = fmaxf(1.E-10f, qv1d);
        float rho = 0.622f * pres / (KP_R * temp * (qv + 0.622f));
        float rc, nc, ri, ni, rr, nr, rs, rg;
        bool L_qc, L_qi, L_qr, L_qs, L_qg;
        if (qc1d > R1) {
          rc = qc1d * rho;
          L_qc = true;
          nc = Nt_c;   // the lamc/xDc clamps of M:1399-1408 only feed nc, overwritten at M:1410
        } else {
          qc1d = 0.0f; rc = R1; nc = 2.f; L_qc = false;
        }
        if (qi1d > R1) {
          ri = qi1d * rho;
          ni = fmaxf(R2, ni1d * rho);
          if (ni <= R2) {
            lami = (double)(ck.cie[1] / 25.E-6f);
            ni = (float)fmin(499.E3, (double)(ck.cig[0] * ck.oig2 * ri / ck.am_i) * cube_d(lami));
          }
          L_qi = true;
          lami = ice_lam(ni, ri);
          ilami = (double)1.f / lami;
          const float xDi = (float)((double)(3.f + 0.f + 1.f) * ilami);
          if (xDi < 5.E-6f) {
            const double l2 = (double)(ck.cie[1] / 5.E-6f);
            ni = (float)fmin(499.E3, (double)(ck.cig[0] * ck.oig2 * ri / ck.am_i) * cube_d(l2));
            lami_stale = true;
          } else if (xDi > 300.E-6f) {
            const double l2 = (double)(ck.cie[1] / 300.E-6f);
            ni = (float)((double)(ck.cig[0] * ck.oig2 * ri / ck.am_i) * cube_d(l2));
            lami_stale = true;
          }
        } else {
          qi1d = 0.0f; ni1d = 0.0f; ri = R1; ni = R2; L_qi = false;
        }
        if (qr1d > R1) {
          rr = qr1d * rho;
          nr = fmaxf(R2, nr1d * rho);
          if (nr <= R2) { mvd_r = 1.0E-3f; nr = nr_from_mvd(rr, mvd_r); }
          L_qr = true;
          lamr = rain_lam(nr, rr);
          mvd_r = (float)((double)(3.0f + 0.f + 0.672f) / lamr);
          if (mvd_r > 2.5E-3f) { mvd_r = 2.5E-3f; nr = nr_from_mvd(rr, mvd_r); lamr_stale = true; }
          else if (mvd_r < D0r * 0.75f) { mvd_r = D0r * 0.75f; nr = nr_from_mvd(rr, mvd_r); lamr_stale = true; }
        } else {
          qr1d = 0.0f; nr1d = 0.0f; rr = R1; nr = R2; L_qr = false;
        }
        if (qs1d > R1) { rs = qs1d * rho; L_qs = true; } else { qs1d = 0.0f; rs = R1; L_qs = false; }
        if (qg1d > R1) { rg = qg1d * rho; L_qg = true; } else { qg1d = 0.0f; rg = R1; L_qg = false; }

        // the inputs are not needed again before S8: parked in shared memory while the rates fill the registers
        s_in[tid] = t1d; s_in[NT + tid] = qv1d; s_in[2 * NT + tid] = qc1d; s_in[3 * NT + tid] = qi1d; s_in[4 * NT + tid] = qr1d;
        s_in[5 * NT + tid] = qs1d; s_in[6 * NT + tid] = qg1d; s_in[7 * NT + tid] = ni1d; s_in[8 * NT + tid] = nr1d;

        // ---- S2, M:1503-1533 -------------------------------------------------------------------
        float tempc = temp - 273.15f;
        float rhof = sqrtf(ck.rho_not / rho);
        float rhof2 = sqrtf(rhof);
        float qvs = rslf(pres, temp);
        const float qvsi = (tempc <= 0.0f) ? rsif(pres, temp) : qvs;
        float ssatw = qv / qvs - 1.f;
        float ssati = qv / qvsi - 1.f;
        if (fabsf(ssatw) < EPSF) ssatw = 0.0f;
        if (fabsf(ssati) < EPSF) ssati = 0.0f;
        // diffu (M:1512) is read by vapour deposition / sublimation and melting of ice, snow and graupel only (M:1896,
        // M:2126, M:2156, M:2166, M:2238, M:2255); rain evaporation evaluates its own (M:2888)
        const bool ice_any = !iiwarm && (L_qi || L_qs || L_qg);
        float diffu = 0.f;
        if (ice_any) diffu = 2.11E-5f * pow_f(temp / 273.15f, 1.94f) * (101325.f / pres);
        float visco = (tempc >= 0.0f) ? (1.718f + 0.0049f * tempc) * 1.0E-5f
                                      : (1.718f + 0.0049f * tempc - 1.2E-5f * tempc * tempc) * 1.0E-5f;
        float ocp = 1.f / (KP_CP * (1.f + 0.887f * qv));
        float vsc2 = sqrtf(rho / visco);
        float lvap = KP_LVAP0 + (2106.0f - 4218.0f) * tempc;
        float tcond = (5.69f + 0.0168f * tempc) * 1.0E-5f * 418.936f;

        if (!iiwarm) {
          // ---- S3, M:1545-1628 snow moments ----------------------------------------------------
          if (L_qs) {
            const float tc0 = fminf(-0.1f, temp - 273.15f);
            smob = rs * ck.oams;
            const float smo2 = smob;                    // bm_s = 2 branch of M:1553
            const float* sa = c_sa; const float* sb = c_sb;
            // Of the six moments of M:1555-1626 the 1st and the (1+(bv_s+1)/2)-th feed deposition, sublimation and melting
            // of every snow level.  The 0th is only read by the melting number rate (M:2242, at or above 0 C), the
            // (bm_s+1)-th by riming (M:1905, with cloud water) and the (bv_s+2)-th by riming and by the collection of
            // cloud ice (M:1910, M:2185): those three are evaluated where they are read.
            float loga_ = sa[1] + sa[2] * tc0 + sa[3] + sa[4] * tc0 + sa[5] * tc0 * tc0 + sa[6] + sa[7] * tc0 * tc0
                          + sa[8] * tc0 + sa[9] * tc0 * tc0 * tc0 + sa[10];
            float b_ = sb[1] + sb[2] * tc0 + sb[3] + sb[4] * tc0 + sb[5] * tc0 * tc0 + sb[6] + sb[7] * tc0 * tc0
                       + sb[8] * tc0 + sb[9] * tc0 * tc0 * tc0 + sb[10];
            smo1 = pow10_f(loga_) * pow_f(smo2, b_);
            smof = field_moment(tc0, ck.cse[15], smo2);
          }
          // ---- S4, M:1633-1654 graupel intercept ------------------------------------------------
          if (temp >= 270.65f) warm_above_a = true;
          { double nm = N0_min_a; graupel_n0(!warm_above_a && k > 0, L_qr, L_qg, mvd_r, rg, n0_empty, nm, ilamg, N0_g); N0_min_a = (float)nm; }
        }
        // M:1661-1666 rain slope and intercept.  Without rain (rr = R1, nr = R2) every reader of lamr, ilamr, N0_r
        // and mvd_r is switched off (L_qr at M:1676, M:1724, M:2880; rr >= r_r(1) at M:1818, M:1964, M:2028, M:2188)
        if (L_qr) {
          if (lamr_stale) { lamr = rain_lam(nr, rr); mvd_r = (float)((double)(3.0f + 0.f + 0.672f) / lamr); }
          ilamr = (double)1.f / lamr;
          N0_r = (double)(nr * ck.org2) * lamr;                                // lamr**cre(2), cre(2) = 1
        }

        // ---- S5, M:1676-1742 warm rain -----------------------------------------------------------
        if (L_qr && mvd_r > D0r) {
          const float Ef_rr = 1.0f - exp_f(2300.0f * (mvd_r - 1950.0E-6f));
          pnr_rcr = (double)(Ef_rr * 2.0f * nr * rr);
          nr_acc -= pnr_rcr;
        }
        mvd_c = D0c;
        if (L_qc) {
          nu_c = min(15, nint_f(1000.E6f / nc) + 2);
          xDc = fmaxf(D0c * 1.E6f, pow_f(rc / (ck.am_r * nc), ck.obmr) * 1.E6f);
          lamc = (double)pow_f(nc * ck.am_r * ck.ccg[1][nu_c - 1] * ck.ocg1[nu_c - 1] / rc, ck.obmr);
          mvd_c = (float)((double)(3.0f + (float)nu_c + 0.672f) / lamc);
        }
        if (rc > 0.01e-3f) {
          const float Dc_g = (float)(((double)ck.dcg_fac[nu_c - 1] / lamc) * (double)1.E6f);
          const float Dc_b = pow_f(xDc * xDc * xDc * Dc_g * Dc_g * Dc_g - xDc * xDc * xDc * xDc * xDc * xDc, 1.f / 6.f);
          const float zeta1 = 0.5f * ((6.25E-6f * xDc * Dc_b * Dc_b * Dc_b - 0.4f) + fabsf(6.25E-6f * xDc * Dc_b * Dc_b * Dc_b - 0.4f));
          const float zeta = 0.027f * rc * zeta1;
          // Below the autoconversion threshold zeta is exactly +0 (most cloudy cells) and so are the three rates: the
          // divisions are skipped there (0/x takes the slow path of the division routines: 3 x ~60 instructions for
          // nearly every cloud level, profiles/r01).  A NaN zeta (negative argument of the 6th root) takes the full path.
          if (!(zeta == 0.0f)) {
            const float taud = 0.5f * ((0.5f * Dc_b - 7.5f) + fabsf(0.5f * Dc_b - 7.5f)) + R1;
            const float tau = 3.72f / (rc * taud);
            prr_wau = (double)(zeta / tau);
            prr_wau = fmin((double)(rc * odts), prr_wau);
            pnr_wau = prr_wau / (double)(ck.am_r * (float)nu_c * D0r * D0r * D0r);
            pnc_wau = fmin((double)(nc * odts), prr_wau / (double)(ck.am_r * mvd_c * mvd_c * mvd_c));
            nr_acc += pnr_wau; nc_acc -= pnc_wau;
          }
        }
        if (L_qr && mvd_r > D0r && mvd_c > D0c) {
          lamr = (double)1.f / ilamr;
          int idx = 1 + (int)((double)NBINS * log((double)mvd_r / ck.Dr1) / ck.lnDr);
          idx = min(idx, (int)NBINS);
          int jc = (int)(mvd_c * 1.E6f);
          jc = max(1, min(jc, (int)NBINS));                                   // U11: bound the unbounded subscript
          const float Ef_rw = ck.efrw[(idx - 1) + NBINS * (jc - 1)];
          const double lf4 = 1.0 / sq_d(sq_d(lamr + (double)KP_FV_R));          // (lamr+fv_r)**(-cre(9)), cre(9) = 4
          prr_rcw = (double)(rhof * ck.t1_qr_qc * Ef_rw * rc) * N0_r * lf4;
          prr_rcw = fmin((double)(rc * odts), prr_rcw);
          pnc_rcw = (double)(rhof * ck.t1_qr_qc * Ef_rw * nc) * N0_r * lf4;
          pnc_rcw = fmin((double)(nc * odts), pnc_rcw);
          nc_acc -= pnc_rcw;
        }

        LOCKBAR(1);
        // ---- S6, M:1749-2286 ice-phase processes --------------------------------------------------
        if (!iiwarm) {
          vts_boost = 1.5f;
          tempc = temp - 273.15f;
          const int idx_tc = max(1, min(nint_f(-tempc), 45));
          int idx_t = (int)((tempc - 2.5f) / 5.f) - 1;
          idx_t = max(1, -idx_t);
          idx_t = min(idx_t, (int)NTB_T);
          const int idx_c = (rc > ck.r_c1) ? decade_idx_f(rc, ck.nic2, NTB_C) : 1;
          const int idx_i = (ri > ck.r_i1) ? decade_idx_f(ri, ck.nii2, NTB_I) : 1;
          const int idx_i1 = (ni > ck.Nt_i1) ? decade_idx_f(ni, ck.nii3, NTB_I1) : 1;
          int idx_r = 1, idx_r1 = NTB_R1, idx_s, idx_g = 1, idx_g1 = NTB_G1;
          if (rr > ck.r_r1) {
            idx_r = decade_idx_f(rr, ck.nir2, NTB_R);
            lamr = (double)1.f / ilamr;
            const double lam_exp = lamr * (double)ck.n0r_fac;
            const double N0_exp = (double)(ck.org1 * rr / ck.am_r) * sq_d(sq_d(lam_exp));   // **cre(1), cre(1) = 4
            idx_r1 = decade_idx_d(N0_exp, ck.nir3, NTB_R1);
          }
          idx_s = (rs > ck.r_s1) ? decade_idx_f(rs, ck.nis2, NTB_S) : 1;
          if (rg > ck.r_g1) {
            idx_g = decade_idx_f(rg, ck.nig2, NTB_G);
            const double lamg = (double)1.f / ilamg;
            const double lam_exp = lamg * (double)ck.n0g_fac;
            const double N0_exp = (double)(ck.ogg1 * rg / ck.am_g) * sq_d(sq_d(lam_exp));   // **cge(1), cge(1) = 4
            idx_g1 = decade_idx_d(N0_exp, ck.nig3, NTB_G1);
          }

          // M:1884-1900 sublimation/deposition prefactor
          // (rvs and t1_subl are read by the deposition / sublimation rates of ice, snow and graupel only)
          float rvs = 0.f, t1_subl = 0.f;
          if (ice_any) {
            const float otemp = 1.f / temp;
            const float lsub = KP_LSUB, oRv = ck.oRv;
            rvs = rho * qvsi;
            const float rvs_p = rvs * otemp * (lsub * otemp * oRv - 1.f);
            const float rvs_pp = rvs * (otemp * (lsub * otemp * oRv - 1.f) * otemp * (lsub * otemp * oRv - 1.f)
                                        + (-2.f * lsub * otemp * otemp * otemp * oRv) + otemp * otemp);
            const float gamsc = lsub * diffu / tcond * rvs_p;
            float alphsc = 0.5f * (gamsc / (1.f + gamsc)) * (gamsc / (1.f + gamsc)) * rvs_pp / rvs_p * rvs / rvs_p;
            alphsc = fmaxf(1.E-9f, alphsc);
            float xsat = ssati;
            if (fabsf(xsat) < 1.E-9f) xsat = 0.f;
            t1_subl = 4.f * KP_PI * (1.0f - alphsc * xsat + 2.f * alphsc * alphsc * xsat * xsat
                                     - 5.f * alphsc * alphsc * alphsc * xsat * xsat * xsat) / (1.f + gamsc);
          }

          // M:1903-1935 riming of snow and graupel
          if (L_qc && mvd_c > D0c) {
            float xDs = 0.0f;
            if (L_qs) { smoc = field_moment(fminf(-0.1f, temp - 273.15f), ck.cse[0], smob); xDs = smoc / smob; }
            if (xDs > D0s) {
              smoe = field_moment(fminf(-0.1f, temp - 273.15f), ck.cse[12], smob); have_smoe = true;
              int idx = 1 + (int)((double)NBINS * log((double)xDs / ck.Ds1) / ck.lnDs);
              idx = min(idx, (int)NBINS);
              int jc = (int)(mvd_c * 1.E6f);
              jc = max(1, min(jc, (int)NBINS));                               // U11
              const float Ef_sw = ck.efsw[(idx - 1) + NBINS * (jc - 1)];
              prs_scw = (double)(rhof * ck.t1_qs_qc * Ef_sw * rc * smoe);
              pnc_scw = (double)(rhof * ck.t1_qs_qc * Ef_sw * nc * smoe);
              pnc_scw = fmin((double)(nc * odts), pnc_scw);
              nc_acc -= pnc_scw;
            }
            if (rg >= ck.r_g1 && mvd_c > D0c) {
              const float xDg = (float)((double)(3.f + 0.f + 1.f) * ilamg);
              const float vtg = (float)((double)(rhof * KP_AV_G * ck.cgg[5] * ck.ogg3) * pow_d(ilamg, (double)KP_BV_G));
              const float stoke_g = mvd_c * mvd_c * vtg * KP_RHO_W / (9.f * visco * xDg);
              if (xDg > D0g) {
                float Ef_gw = 0.0f;
                if (stoke_g >= 0.4f && stoke_g <= 10.f) Ef_gw = 0.55f * log10_f(2.51f * stoke_g);
                else if (stoke_g < 0.4f) Ef_gw = 0.0f;
                else if (stoke_g > 10.f) Ef_gw = 0.77f;
                const double il9 = pow_d(ilamg, (double)ck.cge[8]);
                prg_gcw = (double)(rhof * ck.t1_qg_qc * Ef_gw * rc) * N0_g * il9;
                pnc_gcw = (double)(rhof * ck.t1_qg_qc * Ef_gw * nc) * N0_g * il9;
                pnc_gcw = fmin((double)(nc * odts), pnc_gcw);
                nc_acc -= pnc_gcw;
              }
            }
          }

          // M:1964-2019 rain-snow and rain-graupel collection tables (interleaved records)
          if (rr >= ck.r_r1) {
            if (rs >= ck.r_s1) {
              const double* rec = ck.racs + ((size_t)(idx_s - 1) + (size_t)NTB_S * ((idx_t - 1) + (size_t)NTB_T * ((idx_r1 - 1) + (size_t)NTB_R1 * (idx_r - 1)))) * S_N;
              const double tmr2 = rec[S_TMR_RACS2], tcr2 = rec[S_TCR_SACR2], tmr1 = rec[S_TMR_RACS1], tcr1 = rec[S_TCR_SACR1];
              const double tcs1 = rec[S_TCS_RACS1], tms1 = rec[S_TMS_SACR1];
              if (temp < T_0) {
                prr_rcs = -(tmr2 + tcr2 + tmr1 + tcr1);
                prs_rcs = tmr2 + tcr2 - tcs1 - tms1;
                prg_rcs = tmr1 + tcr1 + tcs1 + tms1;
                prr_rcs = fmax((double)(-rr * odts), prr_rcs);
                prs_rcs = fmax((double)(-rs * odts), prs_rcs);
                prg_rcs = fmin((double)((rr + rs) * odts), prg_rcs);
                pnr_rcs = rec[S_TNR_RACS1] + rec[S_TNR_RACS2] + rec[S_TNR_SACR1] + rec[S_TNR_SACR2];
              } else {
                prs_rcs = -tcs1 - tms1 + tmr2 + tcr2;
                prs_rcs = fmax((double)(-rs * odts), prs_rcs);
                prr_rcs = -prs_rcs;
                pnr_rcs = rec[S_TNR_RACS2] + rec[S_TNR_SACR2];
              }
              pnr_rcs = fmin((double)(nr * odts), pnr_rcs);
              nr_acc -= pnr_rcs;
            }
            if (rg >= ck.r_g1) {
              const double* rec = ck.racg + ((size_t)(idx_g1 - 1) + (size_t)NTB_G1 * ((idx_g - 1) + (size_t)NTB_G * ((idx_r1 - 1) + (size_t)NTB_R1 * (idx_r - 1)))) * G_N;
              if (temp < T_0) {
                prg_rcg = rec[G_TMR_RACG] + rec[G_TCR_GACR];
                prg_rcg = fmin((double)(rr * odts), prg_rcg);
                prr_rcg = -prg_rcg;
                pnr_rcg = rec[G_TNR_RACG] + rec[G_TNR_GACR];
                pnr_rcg = fmin((double)(nr * odts), pnr_rcg);
                nr_acc -= pnr_rcg;
              } else {
                prr_rcg = rec[G_TCG_RACG];
                prr_rcg = fmin((double)(rg * odts), prr_rcg);
                prg_rcg = -prr_rcg;
                pnr_rcg = (double)-5.f * rec[G_TNR_GACR];
                nr_acc -= pnr_rcg;
              }
            }
          }

          if (temp < T_0) {
            // ---- below freezing, M:2025-2231 ---------------------------------------------------
            vts_boost = 1.0f;
            const float rate_max = (qv - qvsi) * rho * odts * 0.999f;
            if (rr > ck.r_r1) {
              const double* rec = ck.qrfz + ((size_t)(idx_r - 1) + (size_t)NTB_R * ((idx_r1 - 1) + (size_t)NTB_R1 * (idx_tc - 1))) * F_N;
              prg_rfz = rec[F_TPG] * (double)odts;
              pri_rfz = rec[F_TPI] * (double)odts;
              pni_rfz = rec[F_TNI] * (double)odts;
              pnr_rfz = rec[F_TNR] * (double)odts;
              pnr_rfz = fmin((double)(nr * odts), pnr_rfz);
              nr_acc -= pnr_rfz; ni_acc += pni_rfz;
            } else if (rr > R1 && temp < KP_HGFR) {
              pri_rfz = (double)(rr * odts);
              pnr_rfz = (double)(nr * odts);
              pni_rfz = pnr_rfz;
              nr_acc -= pnr_rfz; ni_acc += pni_rfz;
            }
            if (rc > ck.r_c1) {
              const double* rec = ck.qcfz + ((size_t)(idx_c - 1) + (size_t)NTB_C * (idx_tc - 1)) * C_N;
              pri_wfz = rec[C_TPI] * (double)odts;
              pri_wfz = fmin((double)(rc * odts), pri_wfz);
              pni_wfz = rec[C_TNI] * (double)odts;
              pni_wfz = fmin(fmin((double)(Nt_c * odts), pri_wfz / (double)(2.f * KP_XM0I)), pni_wfz);
              ni_acc += pni_wfz; nc_acc -= pni_wfz;
            } else if (rc > R1 && temp < KP_HGFR) {
              pri_wfz = (double)(rc * odts);
              pni_wfz = (double)(nc * odts);
              ni_acc += pni_wfz; nc_acc -= pni_wfz;
            }
            // M:2090-2101 Cooper nucleation
            if ((ssati >= 0.25f) || (ssatw > EPSF && temp < 253.15f)) {
              const float xnc = fminf(250.E3f, KP_TNO * exp_f(KP_ATO * (T_0 - temp)));
              const float xni = (float)((double)ni + (pni_rfz + pni_wfz) * (double)DT);
              pni_inu = (double)(0.5f * (xnc - xni + fabsf(xnc - xni)) * odts);
              pri_inu = fmin((double)rate_max, (double)KP_XM0I * pni_inu);
              pni_inu = (pri_inu == 0.0) ? pri_inu : pri_inu / (double)KP_XM0I;      // (a zero keeps its sign either way)
              ni_acc += pni_inu;
            }
            // M:2116-2149 deposition / sublimation of cloud ice, ice -> snow
            float oxmi = 0.f, xDi = 0.f;
            if (L_qi) {
              if (lami_stale) lami = ice_lam(ni, ri);
              ilami = (double)1.f / lami;
              xDi = (float)fmax((double)ck.D0i, (double)(3.f + 0.f + 1.f) * ilami);
              const float xmi = ck.am_i * cube_f(xDi);
              oxmi = 1.f / xmi;
              pri_ide = (double)(KP_C_CUBE * t1_subl * diffu * ssati * rvs * ck.oig1 * ck.cig[4] * ni) * ilami;
              const double* rec = ck.iaus + ((size_t)(idx_i - 1) + (size_t)NTB_I * (idx_i1 - 1)) * I_N;
              if (pri_ide < 0.0) {
                pri_ide = fmax(fmax((double)(-ri * odts), pri_ide), (double)rate_max);
                pni_ide = pri_ide * (double)oxmi;
                pni_ide = fmax((double)(-ni * odts), pni_ide);
              } else {
                pri_ide = fmin(pri_ide, (double)rate_max);
                prs_ide = (1.0 - rec[I_TPI_IDE]) * pri_ide;
                pri_ide = rec[I_TPI_IDE] * pri_ide;
              }
              if ((idx_i == NTB_I) || (xDi > 5.0f * D0s)) {
                prs_iau = (double)(ri * .99f * odts);
                pni_iau = (double)(ni * .95f * odts);
              } else if (xDi < 0.1f * D0s) {
                prs_iau = 0.; pni_iau = 0.;
              } else {
                prs_iau = rec[I_TPS] * (double)odts;
                prs_iau = fmin((double)(ri * .99f * odts), prs_iau);
                pni_iau = rec[I_TNI] * (double)odts;
                pni_iau = fmin((double)(ni * .95f * odts), pni_iau);
              }
              ni_acc -= pni_iau;
            }
            // M:2153-2175 deposition / sublimation of snow, sublimation of graupel
            if (L_qs) {
              float C_snow = KP_C_SQRD + (tempc + 1.5f) * (KP_C_CUBE - KP_C_SQRD) / (-30.f + 1.5f);
              C_snow = fmaxf(KP_C_SQRD, fminf(C_snow, KP_C_CUBE));
              prs_sde = (double)(C_snow * t1_subl * diffu * ssati * rvs
                                 * (ck.t1_qs_sd * smo1 + ck.t2_qs_sd * rhof2 * vsc2 * smof));
              if (prs_sde < 0.) prs_sde = fmax(fmax((double)(-rs * odts), prs_sde), (double)rate_max);
              else prs_sde = fmin(prs_sde, (double)rate_max);
            }
            if (L_qg && ssati < -EPSF) {
              prg_gde = (double)(KP_C_CUBE * t1_subl * diffu * ssati * rvs) * N0_g
                        * ((double)ck.t1_qg_sd * sq_d(ilamg)
                           + (double)(ck.t2_qg_sd * vsc2 * rhof2) * pow_d(ilamg, (double)ck.cge[10]));
              if (prg_gde < 0.) prg_gde = fmax(fmax((double)(-rg * odts), prg_gde), (double)rate_max);
              else prg_gde = fmin(prg_gde, (double)rate_max);
            }
            // M:2178-2202 snow and rain collecting cloud ice (lami/xDi/oxmi as recomputed at M:2179-2183)
            if (L_qi) {
              if (rs >= ck.r_s1) {
                if (!have_smoe) smoe = field_moment(fminf(-0.1f, temp - 273.15f), ck.cse[12], smob);
                prs_sci = (double)(ck.t1_qs_qi * rhof * KP_EF_SI * ri * smoe);
                pni_sci = prs_sci * (double)oxmi;
                ni_acc -= pni_sci;
              }
              if (rr >= ck.r_r1 && mvd_r > 4.f * xDi) {
                lamr = (double)1.f / ilamr;
                const double lf = lamr + (double)KP_FV_R;
                const double lf2 = lf * lf, lf4 = 1.0 / (lf2 * lf2), lf7 = 1.0 / (lf2 * lf2 * lf2 * lf);
                pri_rci = (double)(rhof * ck.t1_qr_qi * KP_EF_RI * ri) * N0_r * lf4;
                pnr_rci = (double)(rhof * ck.t1_qr_qi * KP_EF_RI * ni) * N0_r * lf4;
                pni_rci = pri_rci * (double)oxmi;
                nr_acc -= pnr_rci; ni_acc -= pni_rci;
                prr_rci = (double)(rhof * ck.t2_qr_qi * KP_EF_RI * ni) * N0_r * lf7;       // cre(8) = 7
                prr_rci = fmin((double)(rr * odts), prr_rci);
                prg_rci = pri_rci + prr_rci;
              }
            }
            // M:2205-2218 Hallett-Mossop
            if (prg_gcw > (double)EPSF && tempc > -8.0f) {
              float tf = 0.f;
              if (tempc >= -5.0f && tempc < -3.0f) tf = 0.5f * (-3.0f - tempc);
              else if (tempc > -8.0f && tempc < -5.0f) tf = 0.33333333f * (8.0f + tempc);
              pni_ihm = (double)(3.5E8f * tf) * prg_gcw;
              pri_ihm = (double)KP_XM0I * pni_ihm;
              ni_acc += pni_ihm;
              prs_ihm = prs_scw / (prs_scw + prg_gcw) * pri_ihm;
              prg_ihm = prg_gcw / (prs_scw + prg_gcw) * pri_ihm;
            }
            // M:2224-2231 rimed snow -> graupel
            if (prs_scw > (double)2.0f * prs_sde && prs_sde > (double)EPSF) {
              const float r_frac = (float)fmin(30.0, prs_scw / prs_sde);
              const float g_frac = fminf(0.95f, 0.15f + (r_frac - 2.f) * .028f);
              vts_boost = fminf(1.5f, 1.1f + (r_frac - 2.f) * .016f);
              prg_scw = (double)g_frac * prs_scw;
              prs_scw = (double)(1.f - g_frac) * prs_scw;
            }
          } else {
            // ---- at or above freezing, M:2237-2281 ----------------------------------------------
            float delQvs = 0.f;                                                // M:1508, read by the melting terms only
            if (L_qs || L_qg) delQvs = fmaxf(0.0f, rslf(pres, 273.15f) - qv);
            if (L_qs) {
              prr_sml = (double)((tempc * tcond - KP_LVAP0 * diffu * delQvs)
                                 * (ck.t1_qs_me * smo1 + ck.t2_qs_me * rhof2 * vsc2 * smof));
              prr_sml = prr_sml + (double)(4218.f * ck.olfus * tempc) * (prr_rcs + prs_scw);
              prr_sml = fmin((double)(rs * odts), fmax(0., prr_sml));
              {                                          // 0th moment, M:1557-1560
                const float tc0 = fminf(-0.1f, temp - 273.15f);
                const float* sa = c_sa; const float* sb = c_sb;
                const float loga_ = sa[1] + sa[2] * tc0 + sa[5] * tc0 * tc0 + sa[9] * tc0 * tc0 * tc0;
                const float b_ = sb[1] + sb[2] * tc0 + sb[5] * tc0 * tc0 + sb[9] * tc0 * tc0 * tc0;
                smo0 = pow10_f(loga_) * pow_f(smob, b_);
              }
              pnr_sml = (double)(smo0 / rs) * prr_sml * (double)pow10_f(-0.25f * tempc);
              pnr_sml = fmin((double)(smo0 * odts), pnr_sml);
              nr_acc += pnr_sml;
              if (ssati < 0.f) {
                prs_sde = (double)(KP_C_CUBE * t1_subl * diffu * ssati * rvs
                                   * (ck.t1_qs_sd * smo1 + ck.t2_qs_sd * rhof2 * vsc2 * smof));
                prs_sde = fmax((double)(-rs * odts), prs_sde);
              }
            }
            if (L_qg) {
              const double il10 = sq_d(ilamg), il11 = pow_d(ilamg, (double)ck.cge[10]);
              prr_gml = (double)(tempc * tcond - KP_LVAP0 * diffu * delQvs) * N0_g
                        * ((double)ck.t1_qg_me * il10 + (double)(ck.t2_qg_me * rhof2 * vsc2) * il11);
              prr_gml = fmin((double)(rg * odts), fmax(0., prr_gml));
              pnr_gml = N0_g * (double)ck.cgg[1] * ilamg / (double)rg * prr_gml * (double)pow10_f(-0.5f * tempc);
              nr_acc += pnr_gml;
              if (ssati < 0.f) {
                prg_gde = (double)(KP_C_CUBE * t1_subl * diffu * ssati * rvs) * N0_g
                          * ((double)ck.t1_qg_sd * il10 + (double)(ck.t2_qg_sd * vsc2 * rhof2) * il11);
                prg_gde = fmax((double)(-rg * odts), prg_gde);
              }
            }
            if (DT > 120.f) {
              prr_rcw = prr_rcw + prs_scw + prg_gcw;
              prs_scw = 0.; prg_gcw = 0.;
            }
          }
        }

        LOCKBAR(2);
        // ---- S7, M:2291-2387 conservation limiters -----------------------------------------------
        {
          float sump = (float)(pri_inu + pri_ide + prs_ide + prs_sde + prg_gde + 0.0);
          float rate_max = (qv - qvsi) * odts * 0.999f;                          // U7: no rho factor here
          if ((sump > EPSF && sump > rate_max) || (sump < -EPSF && sump < rate_max)) {
            const double ratio = (double)(rate_max / sump);
            pri_inu *= ratio; pri_ide *= ratio; pni_ide *= ratio; prs_ide *= ratio; prs_sde *= ratio; prg_gde *= ratio;
          }
          sump = (float)(-prr_wau - pri_wfz - prr_rcw - prs_scw - prg_scw - prg_gcw);
          rate_max = -rc * odts;
          if (sump < rate_max && L_qc) {
            const double ratio = (double)(rate_max / sump);
            prr_wau *= ratio; pri_wfz *= ratio; prr_rcw *= ratio; prs_scw *= ratio; prg_scw *= ratio; prg_gcw *= ratio;
          }
          sump = (float)(pri_ide - prs_iau - prs_sci - pri_rci);
          rate_max = -ri * odts;
          if (sump < rate_max && L_qi) {
            const double ratio = (double)(rate_max / sump);
            pri_ide *= ratio; prs_iau *= ratio; prs_sci *= ratio; pri_rci *= ratio;
          }
          sump = (float)(-prg_rfz - pri_rfz - prr_rci + prr_rcs + prr_rcg);
          rate_max = -rr * odts;
          if (sump < rate_max && L_qr) {
            const double ratio = (double)(rate_max / sump);
            prg_rfz *= ratio; pri_rfz *= ratio; prr_rci *= ratio; prr_rcs *= ratio; prr_rcg *= ratio;
          }
          sump = (float)(prs_sde - prs_ihm - prr_sml + prs_rcs);
          rate_max = -rs * odts;
          if (sump < rate_max && L_qs) {
            const double ratio = (double)(rate_max / sump);
            prs_sde *= ratio; prs_ihm *= ratio; prr_sml *= ratio; prs_rcs *= ratio;
          }
          sump = (float)(prg_gde - prg_ihm - prr_gml + prg_rcg);
          rate_max = -rg * odts;
          if (sump < rate_max && L_qg) {
            const double ratio = (double)(rate_max / sump);
            prg_gde *= ratio; prg_ihm *= ratio; prr_gml *= ratio; prg_rcg *= ratio;
          }
          pri_ihm = prs_ihm + prg_ihm;
          float ratio = (float)fmin(fabs(prr_rcg), fabs(prg_rcg));
          prr_rcg = (double)(ratio * copysignf(1.0f, (float)prr_rcg));
          prg_rcg = -prr_rcg;
          if (temp > T_0) {
            ratio = (float)fmin(fabs(prr_rcs), fabs(prs_rcs));
            prr_rcs = (double)(ratio * copysignf(1.0f, (float)prr_rcs));
            prs_rcs = -prr_rcs;
          }
        }

        {   // inputs back from shared memory (these names shadow the ones loaded at the top of the level)
        const float t1d = s_in[tid], qv1d = s_in[NT + tid], qc1d = s_in[2 * NT + tid], qi1d = s_in[3 * NT + tid],
                    qr1d = s_in[4 * NT + tid], qs1d = s_in[5 * NT + tid], qg1d = s_in[6 * NT + tid], ni1d = s_in[7 * NT + tid],
                    nr1d = s_in[8 * NT + tid];
        // U1 again (nc1d = 0 without cloud water, M:1409); dz of the level
        const float nc1d = (qc1d > R1) ? Nt_c / (0.622f * pres / (KP_R * t1d * (qv1d + 0.622f))) : 0.0f;
        const float dzq = a.dz_col ? a.dz_col[o + col] : a.dz[k];
        // ---- S8, M:2393-2569 tendencies and number/mass balances ------------------------------------
        float tt, qvt, qct, qit, qrt, qst, qgt, nit, nrt, nct;
        {
          const float orho = 1.f / rho;
          const float lfus2 = KP_LSUB - lvap;
          qvt = (float)((-pri_inu - pri_ide - prs_ide - prs_sde - prg_gde) * (double)orho);
          qct = (float)((-prr_wau - pri_wfz - prr_rcw - prs_scw - prg_scw - prg_gcw) * (double)orho);
          nct = (float)(nc_acc * (double)orho);
          float xrc = fmaxf(R1, (qc1d + qct * DT) * rho);
          float xnc = fmaxf(2.f, (nc1d + nct * DT) * rho);
          if (xrc > R1) {
            const int nu = min(15, nint_f(1000.E6f / xnc) + 2);
            const double lc = (double)pow_f(xnc * ck.am_r * ck.ccg[1][nu - 1] * ck.ocg1[nu - 1] / rc, ck.obmr);
            const float xD = (float)((double)(3.f + (float)nu + 1.f) / lc);
            if (xD < D0c) {
              const double l2 = (double)(ck.cce[1][nu - 1] / D0c);
              xnc = (float)((double)(ck.ccg[0][nu - 1] * ck.ocg2[nu - 1] * xrc / ck.am_r) * cube_d(l2));
              nct = (xnc - nc1d * rho) * odts * orho;
            } else if (xD > D0r * 2.f) {
              const double l2 = (double)(ck.cce[1][nu - 1] / (D0r * 2.f));
              xnc = (float)((double)(ck.ccg[0][nu - 1] * ck.ocg2[nu - 1] * xrc / ck.am_r) * cube_d(l2));
              nct = (xnc - nc1d * rho) * odts * orho;
            }
          } else {
            nct = -nc1d * odts;
          }
          xnc = fmaxf(0.f, (nc1d + nct * DT) * rho);
          if (xnc > KP_NT_C_MAX) nct = (KP_NT_C_MAX - nc1d * rho) * odts * orho;

          qit = (float)((pri_inu + pri_ihm + pri_wfz + pri_rfz + pri_ide - prs_iau - prs_sci - pri_rci) * (double)orho);
          nit = (float)((ni_acc + pni_ide) * (double)orho);
          const float xri = fmaxf(R1, (qi1d + qit * DT) * rho);
          float xni = fmaxf(R2, (ni1d + nit * DT) * rho);
          if (xri > R1) {
            lami = ice_lam(xni, xri);
            ilami = (double)1.f / lami;
            const float xD = (float)((double)(3.f + 0.f + 1.f) * ilami);
            if (xD < 5.E-6f) {
              lami = (double)(ck.cie[1] / 5.E-6f);
              xni = (float)fmin(499.E3, (double)(ck.cig[0] * ck.oig2 * xri / ck.am_i) * cube_d(lami));
              nit = (xni - ni1d * rho) * odts * orho;
            } else if (xD > 300.E-6f) {
              lami = (double)(ck.cie[1] / 300.E-6f);
              xni = (float)((double)(ck.cig[0] * ck.oig2 * xri / ck.am_i) * cube_d(lami));
              nit = (xni - ni1d * rho) * odts * orho;
            }
          } else {
            nit = -ni1d * odts;
          }
          xni = fmaxf(0.f, (ni1d + nit * DT) * rho);
          if (xni > 499.E3f) nit = (499.E3f - ni1d * rho) * odts * orho;

          qrt = (float)((prr_wau + prr_rcw + prr_sml + prr_gml + prr_rcs + prr_rcg - prg_rfz - pri_rfz - prr_rci) * (double)orho);
          nrt = (float)(nr_acc * (double)orho);
          const float xrr = fmaxf(R1, (qr1d + qrt * DT) * rho);
          float xnr = fmaxf(R2, (nr1d + nrt * DT) * rho);
          if (xrr > R1) {
            lamr = rain_lam(xnr, xrr);
            mvd_r = (float)((double)(3.0f + 0.f + 0.672f) / lamr);
            if (mvd_r > 2.5E-3f) {
              mvd_r = 2.5E-3f;
              xnr = nr_from_mvd(xrr, mvd_r);
              nrt = (xnr - nr1d * rho) * odts * orho;
            } else if (mvd_r < D0r * 0.75f) {
              mvd_r = D0r * 0.75f;
              xnr = nr_from_mvd(xrr, mvd_r);
              nrt = (xnr - nr1d * rho) * odts * orho;
            }
          } else {
            qrt = -qr1d * odts;
            nrt = -nr1d * odts;
          }
          qst = (float)((prs_iau + prs_sde + prs_sci + prs_scw + prs_rcs + prs_ide - prs_ihm - prr_sml) * (double)orho);
          qgt = (float)((prg_scw + prg_rfz + prg_gde + prg_rcg + prg_gcw + prg_rci + prg_rcs - prg_ihm - prr_gml) * (double)orho);
          if (temp < T_0) {
            tt = (float)(((double)(KP_LSUB * ocp) * (pri_inu + pri_ide + prs_ide + prs_sde + prg_gde + 0.0)
                          + (double)(lfus2 * ocp) * (pri_wfz + pri_rfz + prg_rfz + prs_scw + prg_scw + prg_gcw + prg_rcs
                                                     + prs_rcs + prr_rci + prg_rcg))
                         * (double)orho * (double)1);
          } else {
            tt = (float)(((double)(ck.lfus * ocp) * (-prr_sml - prr_gml - prr_rcg - prr_rcs)
                          + (double)(KP_LSUB * ocp) * (prs_sde + prg_gde))
                         * (double)orho * (double)1);
          }
        }

        LOCKBAR(3);
        // ---- S9, M:2574-2656 state at tau+1 -------------------------------------------------------
        float lvt2;
        {
          temp = t1d + DT * tt;
          const float otemp = 1.f / temp;
          tempc = temp - 273.15f;
          qv = fmaxf(1.E-10f, qv1d + DT * qvt);
          rho = 0.622f * pres / (KP_R * temp * (qv + 0.622f));
          qvs = rslf(pres, temp);
          ssatw = qv / qvs - 1.f;
          if (fabsf(ssatw) < EPSF) ssatw = 0.0f;
          // rhof, rhof2, diffu, visco, vsc2, tcond of M:2588-2600 are read by rain evaporation (which evaluates them
          // again from the post-condensation state here, S12) and the fall speeds (rhof, S13) only
          lvap = KP_LVAP0 + (2106.0f - 4218.0f) * tempc;
          ocp = 1.f / (KP_CP * (1.f + 0.887f * qv));
          lvt2 = lvap * lvap * ocp * ck.oRv * otemp * otemp;

          if ((qc1d + qct * DT) > R1) { rc = (qc1d + qct * DT) * rho; nc = Nt_c; L_qc = true; }
          else { rc = R1; nc = 2.f; L_qc = false; }
          if ((qi1d + qit * DT) > R1) { ri = (qi1d + qit * DT) * rho; ni = fmaxf(R2, (ni1d + nit * DT) * rho); L_qi = true; }
          else { ri = R1; ni = R2; L_qi = false; }
          if ((qr1d + qrt * DT) > R1) {
            rr = (qr1d + qrt * DT) * rho;
            nr = fmaxf(R2, (nr1d + nrt * DT) * rho);
            L_qr = true;
            lamr = rain_lam(nr, rr);
            lamr_stale = false;
            mvd_r = (float)((double)(3.0f + 0.f + 0.672f) / lamr);
            if (mvd_r > 2.5E-3f) { mvd_r = 2.5E-3f; nr = nr_from_mvd(rr, mvd_r); lamr_stale = true; }
            else if (mvd_r < D0r * 0.75f) { mvd_r = D0r * 0.75f; nr = nr_from_mvd(rr, mvd_r); lamr_stale = true; }
          } else { rr = R1; nr = R2; L_qr = false; }
          if ((qs1d + qst * DT) > R1) { rs = (qs1d + qst * DT) * rho; L_qs = true; } else { rs = R1; L_qs = false; }
          if ((qg1d + qgt * DT) > R1) { rg = (qg1d + qgt * DT) * rho; L_qg = true; } else { rg = R1; L_qg = false; }
        }

        // ---- S10, M:2662-2750 snow moments and intercepts again -------------------------------------
        if (!iiwarm) {
          if (L_qs) {
            const float tc0 = fminf(-0.1f, temp - 273.15f);
            smob = rs * ck.oams;
            smoc = field_moment(tc0, ck.cse[0], smob);
            // smod (M:2706-2717) is not read again by any live code
          }
          if (temp >= 270.65f) warm_above_b = true;
          { double nm = N0_min_b; graupel_n0(!warm_above_b && k > 0, L_qr, L_qg, mvd_r, rg, n0_empty, nm, ilamg, N0_g); N0_min_b = (float)nm; }
        }
        if (L_qr) {                                                             // M:2750-2755, as at M:1661
          if (lamr_stale) { lamr = rain_lam(nr, rr); mvd_r = (float)((double)(3.0f + 0.f + 0.672f) / lamr); lamr_stale = false; }
          ilamr = (double)1.f / lamr;
          N0_r = (double)(nr * ck.org2) * lamr;
        }

        LOCKBAR(4);
        // ---- S11, M:2780-2874 cloud condensation / evaporation ---------------------------------------
        if ((ssatw > EPSF) || (ssatw < -EPSF && L_qc)) {
          const float orho = 1.f / rho;
          float clap = (qv - qvs) / (1.f + lvt2 * qvs);
#pragma unroll
          for (int n = 0; n < 3; ++n) {
            const float e = exp_f(lvt2 * clap);
            const float fcd = qvs * e - qv + clap;
            const float dfcd = qvs * lvt2 * e + 1.f;
            clap = clap - fcd / dfcd;
          }
          const float xrc = rc + clap * rho;
          if (xrc > R1) {
            prw_vcd = (double)(clap * odt);
            if (clap > EPSF) {
              const float xnc = Nt_c;
              pnc_wcd = (double)(0.5f * (xnc - nc + fabsf(xnc - nc)) * odts * orho);
            }
          } else {
            prw_vcd = (double)(-rc * orho * odt);
            pnc_wcd = (double)(-nc * orho * odt);
          }
          qvt = (float)((double)qvt - prw_vcd);
          qct = (float)((double)qct + prw_vcd);
          nct = (float)((double)nct + pnc_wcd);
          tt = (float)((double)tt + (double)(lvap * ocp) * prw_vcd * (double)1);
          rc = fmaxf(R1, (qc1d + DT * qct) * rho);
          nc = Nt_c;
          qv = fmaxf(1.E-10f, qv1d + DT * qvt);
          temp = t1d + DT * tt;
          rho = 0.622f * pres / (KP_R * temp * (qv + 0.622f));
          qvs = rslf(pres, temp);
          ssatw = qv / qvs - 1.f;
        }

        // ---- S12, M:2880-2960 rain evaporation -------------------------------------------------------
        if ((ssatw < -EPSF) && L_qr && (!(prw_vcd > 0.))) {
          tempc = temp - 273.15f;
          const float otemp = 1.f / temp;
          const float orho = 1.f / rho;
          rhof = sqrtf(ck.rho_not * orho);
          rhof2 = sqrtf(rhof);
          diffu = 2.11E-5f * pow_f(temp / 273.15f, 1.94f) * (101325.f / pres);
          visco = (tempc >= 0.0f) ? (1.718f + 0.0049f * tempc) * 1.0E-5f
                                  : (1.718f + 0.0049f * tempc - 1.2E-5f * tempc * tempc) * 1.0E-5f;
          vsc2 = sqrtf(rho / visco);
          lvap = KP_LVAP0 + (2106.0f - 4218.0f) * tempc;
          tcond = (5.69f + 0.0168f * tempc) * 1.0E-5f * 418.936f;
          ocp = 1.f / (KP_CP * (1.f + 0.887f * qv));
          const float oRv = ck.oRv;
          const float rvs = rho * qvs;
          const float rvs_p = rvs * otemp * (lvap * otemp * oRv - 1.f);
          const float rvs_pp = rvs * (otemp * (lvap * otemp * oRv - 1.f) * otemp * (lvap * otemp * oRv - 1.f)
                                      + (-2.f * lvap * otemp * otemp * otemp * oRv) + otemp * otemp);
          const float gamsc = lvap * diffu / tcond * rvs_p;
          float alphsc = 0.5f * (gamsc / (1.f + gamsc)) * (gamsc / (1.f + gamsc)) * rvs_pp / rvs_p * rvs / rvs_p;
          alphsc = fmaxf(1.E-9f, alphsc);
          const float xsat = fminf(-1.E-9f, ssatw);
          const float t1_evap = 2.f * KP_PI * (1.0f - alphsc * xsat + 2.f * alphsc * alphsc * xsat * xsat
                                               - 5.f * alphsc * alphsc * alphsc * xsat * xsat * xsat) / (1.f + gamsc);
          const double lamr_ev = (double)1.f / ilamr;
          if (qv / qvs < 0.95f && rr * orho <= 1.E-8f) {
            prv_rev = (double)(rr * orho * odts);
          } else {
            const double lh = lamr_ev + (double)(0.5f * KP_FV_R);
            prv_rev = (double)(t1_evap * diffu * (-ssatw)) * N0_r * (double)rvs
                      * ((double)ck.t1_qr_ev * sq_d(ilamr)                                     // ilamr**cre(10), = 2
                         + (double)(ck.t2_qr_ev * vsc2 * rhof2) * (1.0 / (lh * lh * lh)));      // **(-cre(11)), = 3
            const float rate_max = fminf((rr * orho * odts), (qvs - qv) * odts);
            prv_rev = fmin((double)rate_max, prv_rev * (double)orho);
            if (prr_gml > 0.0) {
              const float eva_factor = fminf(1.0f, 0.01f + (0.99f - 0.01f) * (tempc / 20.0f));
              prv_rev = prv_rev * (double)eva_factor;
            }
          }
          pnr_rev = fmin((double)(nr * 0.99f * orho * odts), prv_rev * (double)nr / (double)rr);
          qrt = (float)((double)qrt - prv_rev);
          qvt = (float)((double)qvt + prv_rev);
          nrt = (float)((double)nrt - pnr_rev);
          tt = (float)((double)tt - (double)(lvap * ocp) * prv_rev * (double)1);
          rr = fmaxf(R1, (qr1d + DT * qrt) * rho);
          qv = fmaxf(1.E-10f, qv1d + DT * qvt);
          nr = fmaxf(R2, (nr1d + DT * nrt) * rho);
          lamr_stale = true;
          temp = t1d + DT * tt;
          rho = 0.622f * pres / (KP_R * temp * (qv + 0.622f));
        }

        // M:2963-3120 the 36 process rates KiD saves with save_dg (optional buffer [36][nz][ncol])
        if (RATES && a.rates && active) {
          float* rp = a.rates + o + col;
          const long st = (long)nz * ncol;
          const double rv[KIDMP_NRATES] = {pri_inu, pri_ide, prs_ide, prs_sde, prg_gde, pri_wfz, prs_scw, prg_scw, prg_gcw, pri_ihm,
                                           pri_rfz, prs_iau, prs_sci, pri_rci, pni_inu, pni_ihm, pni_wfz, pni_rfz, pni_ide, pni_iau,
                                           pni_sci, pni_rci, prr_sml, prr_gml, pnr_rcs, pnr_rcg, pnr_rci, pnr_sml, pnr_gml, pnr_rfz,
                                           prr_wau, prr_rcw, prv_rev, pnr_wau, pnr_rev, pnr_rcr};
#pragma unroll
          for (int q = 0; q < KIDMP_NRATES; ++q) rp[q * st] = (float)rv[q];
        }

        LOCKBAR(5);
        // ---- S13, M:3206-3354 fall speeds, substep counts (top-down carry) -----------------------------
        rhof = sqrtf(ck.rho_not / rho);
        float v_r, v_nr, v_i = 0.f, v_ni = 0.f, v_s = 0.f, v_g = 0.f;
        if (rr > R1) {
          if (!L_qr || lamr_stale) lamr = rain_lam(nr, rr);                   // M:3227: same nr, rr as at M:2750 otherwise
          const double lf = lamr + (double)KP_FV_R;
          const double l2 = lamr * lamr, lf2 = lf * lf;
          // lamr**cre(3) * (lamr+fv_r)**(-cre(6)), cre(3) = 4, cre(6) = 5
          v_r = (float)((double)(rhof * KP_AV_R * ck.crg[5] * ck.org3) * (l2 * l2) * (1.0 / (lf2 * lf2 * lf)));
          // lamr**cre(12) * (lamr+fv_r)**(-cre(7)), cre(12) = 2.5, cre(7) = 3.5
          v_nr = (float)((double)(rhof * KP_AV_R * ck.crg[6] / ck.crg[11]) * (l2 * sqrt(lamr)) * (1.0 / (lf2 * lf * sqrt(lf))));
        } else {
          v_r = vtr_up; v_nr = vtnr_up;
        }
        if (fmaxf(v_r, v_nr) > 1.E-3f) {
          ksed_r = (unsigned short)max((int)ksed_r, k + 1);
          const float delta_tp = dzq / (fmaxf(v_r, v_nr));
          nstep_r = (unsigned short)min(max((int)nstep_r, (int)(DT / delta_tp + 1.f)), KP_NSTEP_MAX);
        }
        if (!iiwarm) {
          if (ri > R1) {
            lami = ice_lam(ni, ri);
            ilami = (double)1.f / lami;
            v_i = (float)((double)(rhof * KP_AV_I * ck.cig[2] * ck.oig2) * ilami);               // ilami**bv_i, bv_i = 1
            v_ni = (float)((double)(rhof * KP_AV_I * ck.cig[5] / ck.cig[6]) * ilami);
          } else {
            v_i = vti_up; v_ni = vtni_up;
          }
          if (v_i > 1.E-3f) {
            ksed_i = (unsigned short)max((int)ksed_i, k + 1);
            const float delta_tp = dzq / v_i;
            nstep_i = (unsigned short)min(max((int)nstep_i, (int)(DT / delta_tp + 1.f)), KP_NSTEP_MAX);
          }
          if (rs > R1) {
            const float xDs = smoc / smob;
            const float Mrat = 1.f / xDs;
            float ils1 = 1.f / (Mrat * KP_LAM0 + KP_FV_S);
            float ils2 = 1.f / (Mrat * KP_LAM1 + KP_FV_S);
            const float mm = pow_f(Mrat, KP_MU_S);
            const float t1_vts = KP_KAP0 * ck.csg[3] * pow_f(ils1, ck.cse[3]);
            const float t2_vts = KP_KAP1 * mm * ck.csg[9] * pow_f(ils2, ck.cse[9]);
            ils1 = 1.f / (Mrat * KP_LAM0);
            ils2 = 1.f / (Mrat * KP_LAM1);
            const float t3_vts = KP_KAP0 * ck.csg[0] * pow_f(ils1, ck.cse[0]);
            const float t4_vts = KP_KAP1 * mm * ck.csg[6] * pow_f(ils2, ck.cse[6]);
            const float vts = rhof * KP_AV_S * (t1_vts + t2_vts) / (t3_vts + t4_vts);
            if (temp > (T_0 + 0.1f)) v_s = fmaxf(vts * vts_boost, vts * ((v_r - vts * vts_boost) / (temp - T_0)));
            else v_s = vts * vts_boost;
          } else {
            v_s = vts_up;
          }
          if (v_s > 1.E-3f) {
            ksed_s = (unsigned short)max((int)ksed_s, k + 1);
            const float delta_tp = dzq / v_s;
            nstep_s = (unsigned short)min(max((int)nstep_s, (int)(DT / delta_tp + 1.f)), KP_NSTEP_MAX);
          }
          if (rg > R1) {
            const float vtg = (float)((double)(rhof * KP_AV_G * ck.cgg[5] * ck.ogg3) * pow_d(ilamg, (double)KP_BV_G));
            v_g = (temp > T_0) ? fmaxf(vtg, v_r) : vtg;
          } else {
            v_g = vtg_up;
          }
          if (v_g > 1.E-3f) {
            ksed_g = (unsigned short)max((int)ksed_g, k + 1);
            const float delta_tp = dzq / v_g;
            nstep_g = (unsigned short)min(max((int)nstep_g, (int)(DT / delta_tp + 1.f)), KP_NSTEP_MAX);
          }
        }
        vtr_up = v_r; vtnr_up = v_nr; vti_up = v_i; vtni_up = v_ni; vts_up = v_s; vtg_up = v_g;

        // hand-off to the sedimentation kernel: [SC_N][nz][ncol], coalesced fire-and-forget stores.
        // S15 (M:3584-3606) needs lfus*ocp where the level ends above T_0 and lfus2*ocp where it ends
        // below HGFR (never both): one signed value carries the product and the case.
        if (FUSE && active) {
          HandOff h;
          h.tt = tt; h.qvt = qvt; h.qct = qct; h.qit = qit; h.qrt = qrt; h.qst = qst; h.qgt = qgt; h.nit = nit; h.nrt = nrt; h.nct = nct;
          h.rr = rr; h.nr = nr; h.ri = ri; h.ni = ni; h.rs = rs; h.rg = rg;
          h.v_r = v_r; h.v_nr = v_nr; h.v_i = v_i; h.v_ni = v_ni; h.v_s = v_s; h.v_g = v_g; h.rho = rho;
          h.s15 = 0.0f;
          if (temp > T_0) h.s15 = ck.lfus * ocp;
          else if (temp < KP_HGFR) h.s15 = -((KP_LSUB - lvap) * ocp);
          finish(h, k, o, dzq, t1d, qv1d, qc1d, qi1d, qr1d, qs1d, qg1d, ni1d, nr1d, pres);
        } else if (active) {
          float s15 = 0.0f;
          if (temp > T_0) s15 = ck.lfus * ocp;
          else if (temp < KP_HGFR) s15 = -((KP_LSUB - lvap) * ocp);
          float* sc = a.scratch + o + col;
          const long ss = (long)nz * ncol;
          sc[SC_TTEN * ss] = tt; sc[SC_QVTEN * ss] = qvt; sc[SC_QCTEN * ss] = qct; sc[SC_QITEN * ss] = qit;
          sc[SC_QRTEN * ss] = qrt; sc[SC_QSTEN * ss] = qst; sc[SC_QGTEN * ss] = qgt; sc[SC_NITEN * ss] = nit;
          sc[SC_NRTEN * ss] = nrt; sc[SC_NCTEN * ss] = nct;
          sc[SC_RR * ss] = rr; sc[SC_NR * ss] = nr; sc[SC_RI * ss] = ri; sc[SC_NI * ss] = ni; sc[SC_RS * ss] = rs; sc[SC_RG * ss] = rg;
          sc[SC_VTR * ss] = v_r; sc[SC_VTNR * ss] = v_nr; sc[SC_VTI * ss] = v_i; sc[SC_VTNI * ss] = v_ni;
          sc[SC_VTS * ss] = v_s; sc[SC_VTG * ss] = v_g; sc[SC_RHO * ss] = rho; sc[SC_S15 * ss] = s15;
        }
        }   // shadowed inputs
        }
      }

      // column summary for the sedimentation kernel: [8][ncol] substep counts and top sedimenting levels
      if (active) {
        int* ci = a.colint + col;
        // U12: the reference leaves the sub-step count unbounded (M:3242); on non-physical input (dt*v/dz in the
        // millions) that is a kernel that never ends, so it is capped where no real case comes near
        ci[0] = nstep_r; ci[ncol] = nstep_i; ci[2 * ncol] = nstep_s; ci[3 * ncol] = nstep_g;      // capped when they were stored
        ci[4 * ncol] = ksed_r; ci[5 * ncol] = ksed_i; ci[6 * ncol] = ksed_s; ci[7 * ncol] = ksed_g;
        if (max(max((int)nstep_r, (int)nstep_i), max((int)nstep_s, (int)nstep_g)) > 1) {     // counted in both modes: the host picks the mode of the next step
          const int at = atomicAdd(a.redo_count, 1);
          if (FUSE) a.redo_list[at] = (int)col;
        }
      }
    }
  }
}

#undef LOCKBAR
#undef n0_empty
#undef N0_min_a
#undef N0_min_b
#undef vtr_up
#undef vtnr_up
#undef vti_up
#undef vtni_up
#undef vts_up
#undef vtg_up
#undef nstep_r
#undef nstep_i
#undef nstep_s
#undef nstep_g
#undef ksed_r
#undef ksed_i
#undef ksed_s
#undef ksed_g

// ---- inputs of the columns on the redo list back from the hand-off buffer (see FUSE above) ----------------------------
__global__ void __launch_bounds__(32) k_restore(StepArgs a) {
  const int slot = blockIdx.x * 32 + threadIdx.x;
  if (slot >= *a.work_count) return;
  const long col = a.work_list[slot];
  const long ncol = a.ncol, ss = (long)a.nz * ncol;
  for (int k = 0; k < a.nz; ++k) {
    const long g = (long)k * ncol + col;
#pragma unroll
    for (int q = 0; q < KIDMP_NFIELDS; ++q) a.f[q][g] = a.scratch[q * ss + g];
  }
}

// ---- K2: sub-stepped upwind sedimentation (M:3365-3578), instant melt / freeze (M:3584-3606), apply
// tendencies and final clamps (M:3623-3686).  One thread per column, light on registers, so many
// warps per SM hide the latency of streaming the hand-off arrays.  All but the last sub-step of a
// species update the hand-off arrays in place; the last sub-step of all four species is fused with
// S15/S16 and the output stores in one top-down sweep (the common nstep = 1 case is that sweep only).
__device__ __forceinline__ void sed_substeps(float* __restrict__ r, float* __restrict__ rten, const float* __restrict__ v,
                                             float* __restrict__ n, float* __restrict__ nten, const float* __restrict__ vn,
                                             const float* __restrict__ rhoa, const float* __restrict__ dz, long dzs, int nz, long ncol,
                                             int nsub, int ksed, float onstep, float DT, bool on, float nfloor, float& ppt) {
  for (int it = 0; it < nsub; ++it) {
    float sr_up = 0.f, sn_up = 0.f, sr_k = 0.f, r0 = 0.f;
#pragma unroll 1
    for (int k = nz - 1; k >= 0; --k) {
      const long o = (long)k * ncol;
      const float rk = r[o];
      const float sr = on ? v[o] * rk : 0.f;
      const float odzq = 1.f / dz[k * dzs], orho = 1.f / rhoa[o];
      float nk = 0.f, sn = 0.f;
      if (n) { nk = n[o]; sn = on ? vn[o] * nk : 0.f; }
      if (k == nz - 1) {
        rten[o] = rten[o] - sr * odzq * onstep * orho;
        r0 = fmaxf(KP_R1, rk - sr * odzq * DT * onstep);
        r[o] = r0;
        if (n) { nten[o] = nten[o] - sn * odzq * onstep * orho; n[o] = fmaxf(nfloor, nk - sn * odzq * DT * onstep); }
      } else if (k + 1 <= ksed) {
        rten[o] = rten[o] + (sr_up - sr) * odzq * onstep * orho;
        r0 = fmaxf(KP_R1, rk + (sr_up - sr) * odzq * DT * onstep);
        r[o] = r0;
        if (n) { nten[o] = nten[o] + (sn_up - sn) * odzq * onstep * orho; n[o] = fmaxf(nfloor, nk + (sn_up - sn) * odzq * DT * onstep); }
      } else {
        r0 = rk;
      }
      sr_up = sr; sn_up = sn; sr_k = sr;
    }
    if (r0 > KP_R1 * 10.f) ppt = ppt + sr_k * DT * onstep;
  }
}

#ifndef K2_MINB
#define K2_MINB 20         // 96 registers, no spills: 20 one-warp blocks per SM
#endif
__global__ void __launch_bounds__(32, K2_MINB) k_sediment(StepArgs a) {
  const int slot = blockIdx.x * 32 + threadIdx.x;          // cloudy columns only: the compacted work list
  if (slot >= *a.work_count) return;
  const long col = a.work_list[slot];
  const int nz = a.nz;
  const long ncol = a.ncol;
  const float DT = a.dt, odt = 1.f / DT;
  float ppt_r = 0.f, ppt_i = 0.f, ppt_s = 0.f, ppt_g = 0.f;
  double lwp = 0.0, iwp = 0.0;
  {
    const int* ci = a.colint + col;
    const int nstep_r = ci[0];
    {
      const int nstep_i = ci[ncol], nstep_s = ci[2 * ncol], nstep_g = ci[3 * ncol];
      int ksed_r = ci[4 * ncol], ksed_i = ci[5 * ncol], ksed_s = ci[6 * ncol], ksed_g = ci[7 * ncol];
      const int kte = nz;
      if (ksed_r == kte) ksed_r = kte - 1;
      if (ksed_i == kte) ksed_i = kte - 1;
      if (ksed_s == kte) ksed_s = kte - 1;
      if (ksed_g == kte) ksed_g = kte - 1;
      const float on_r = nstep_r > 0 ? 1.f / (float)nstep_r : 1.0f, on_i = nstep_i > 0 ? 1.f / (float)nstep_i : 1.0f;
      const float on_s = nstep_s > 0 ? 1.f / (float)nstep_s : 1.0f, on_g = nstep_g > 0 ? 1.f / (float)nstep_g : 1.0f;
      const int n_r = nint_f(1.f / on_r), n_i = nint_f(1.f / on_i), n_s = nint_f(1.f / on_s), n_g = nint_f(1.f / on_g);
      const bool sedi = ck.l_sediment != 0;
      const long ss = (long)nz * ncol;
      float* sc = a.scratch + col;
      const float* rhoa = sc + SC_RHO * ss;
      // layer depths: one vector shared by all columns (KiD) or this column's own (WRF entry)
      const float* const dzp = a.dz_col ? a.dz_col + col : a.dz;
      const long dzs = a.dz_col ? ncol : 1;
      // all but the last sub-step (rain is never gated by l_sediment, U6; the cloud-water stub M:3414-3425 is a no-op, U2)
      if (n_r > 1) sed_substeps(sc + SC_RR * ss, sc + SC_QRTEN * ss, sc + SC_VTR * ss, sc + SC_NR * ss, sc + SC_NRTEN * ss,
                                sc + SC_VTNR * ss, rhoa, dzp, dzs, nz, ncol, n_r - 1, ksed_r, on_r, DT, true, KP_R2, ppt_r);
      if (n_i > 1) sed_substeps(sc + SC_RI * ss, sc + SC_QITEN * ss, sc + SC_VTI * ss, sc + SC_NI * ss, sc + SC_NITEN * ss,
                                sc + SC_VTNI * ss, rhoa, dzp, dzs, nz, ncol, n_i - 1, ksed_i, on_i, DT, sedi, KP_R2, ppt_i);
      if (n_s > 1) sed_substeps(sc + SC_RS * ss, sc + SC_QSTEN * ss, sc + SC_VTS * ss, nullptr, nullptr, nullptr, rhoa, dzp, dzs,
                                nz, ncol, n_s - 1, ksed_s, on_s, DT, sedi, 0.f, ppt_s);
      if (n_g > 1) sed_substeps(sc + SC_RG * ss, sc + SC_QGTEN * ss, sc + SC_VTG * ss, nullptr, nullptr, nullptr, rhoa, dzp, dzs,
                                nz, ncol, n_g - 1, ksed_g, on_g, DT, sedi, 0.f, ppt_g);

      // last sub-step of every species + S15 + S16, one top-down sweep
      const float* __restrict__ Gp = a.p + col;
      float* Gqv = a.f[F_QV] + col; float* Gqc = a.f[F_QC] + col; float* Gqi = a.f[F_QI] + col;
      float* Gqr = a.f[F_QR] + col; float* Gqs = a.f[F_QS] + col; float* Gqg = a.f[F_QG] + col;
      float* Gni = a.f[F_NI] + col; float* Gnr = a.f[F_NR] + col; float* Gt = a.f[F_T] + col;
      SedParams sp;
      sp.DT = DT; sp.odt = odt; sp.on_r = on_r; sp.on_i = on_i; sp.on_s = on_s; sp.on_g = on_g; sp.Nt_c = ck.Nt_c;
      sp.top_r = ksed_r; sp.top_i = ksed_i; sp.top_s = ksed_s; sp.top_g = ksed_g; sp.sedi = sedi; sp.iiwarm = ck.iiwarm != 0;
      SedCarry c;
      c.sr_up = 0.f; c.snr_up = 0.f; c.si_up = 0.f; c.sni_up = 0.f; c.ss_up = 0.f; c.sg_up = 0.f;
      c.ppt_r = ppt_r; c.ppt_i = ppt_i; c.ppt_s = ppt_s; c.ppt_g = ppt_g; c.lwp = 0.0; c.iwp = 0.0;
#pragma unroll 1
      for (int k = nz - 1; k >= 0; --k) {
        const long o = (long)k * ncol;
        const float* q = sc + o;
        HandOff h;
        h.tt = q[SC_TTEN * ss]; h.qvt = q[SC_QVTEN * ss]; h.qct = q[SC_QCTEN * ss]; h.qit = q[SC_QITEN * ss];
        h.qrt = q[SC_QRTEN * ss]; h.qst = q[SC_QSTEN * ss]; h.qgt = q[SC_QGTEN * ss]; h.nit = q[SC_NITEN * ss];
        h.nrt = q[SC_NRTEN * ss]; h.nct = q[SC_NCTEN * ss];
        h.rho = q[SC_RHO * ss]; h.s15 = q[SC_S15 * ss];
        h.rr = q[SC_RR * ss]; h.nr = q[SC_NR * ss]; h.ri = q[SC_RI * ss]; h.ni = q[SC_NI * ss]; h.rs = q[SC_RS * ss]; h.rg = q[SC_RG * ss];
        h.v_r = q[SC_VTR * ss]; h.v_nr = q[SC_VTNR * ss]; h.v_i = q[SC_VTI * ss]; h.v_ni = q[SC_VTNI * ss];
        h.v_s = q[SC_VTS * ss]; h.v_g = q[SC_VTG * ss];
        finish_level(a, sp, c, h, k, nz, o + col, dzp[k * dzs], Gt[o], Gqv[o], Gqc[o], Gqi[o], Gqr[o], Gqs[o], Gqg[o], Gni[o], Gnr[o], Gp[o]);
      }
      ppt_r = c.ppt_r; ppt_i = c.ppt_i; ppt_s = c.ppt_s; ppt_g = c.ppt_g; lwp = c.lwp; iwp = c.iwp;
    }
    // ppt is overwritten with this step's amounts: rain, ice, snow, graupel (I:55-58, I:162-177)
    a.ppt[col] = ppt_r; a.ppt[ncol + col] = ppt_i; a.ppt[2 * ncol + col] = ppt_s; a.ppt[3 * ncol + col] = ppt_g;
    a.coldiag[col] = lwp; a.coldiag[ncol + col] = iwp;      // summed in column order by k_diag_columns
  }
}

// Domain sums in COLUMN order, whatever order the work list had: bitwise reproducible run to run and across
// different work-list orders.  Fixed grid: block b sums the columns [b*chunk, (b+1)*chunk) (thread-strided, then a
// tree), the per-block partials are added up by k_diag_reduce.
__global__ void __launch_bounds__(256) k_diag_columns(StepArgs a, long chunk) {
  __shared__ double s[256];
  const long c0 = (long)blockIdx.x * chunk, c1 = min(c0 + chunk, a.ncol);
  const long ncol = a.ncol;
  double v[KIDMP_NDIAG] = {0., 0., 0., 0., 0., 0., 0., 0.};
  for (long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    v[7] += 1.0;
    if (a.colint[c] < 0) continue;                          // clear sky: ppt = 0, no condensate
    v[0] += (double)a.ppt[c]; v[1] += (double)a.ppt[ncol + c]; v[2] += (double)a.ppt[2 * ncol + c]; v[3] += (double)a.ppt[3 * ncol + c];
    v[4] += a.coldiag[c]; v[5] += a.coldiag[ncol + c];
    v[6] += 1.0;
  }
  for (int q = 0; q < KIDMP_NDIAG; ++q) {
    s[threadIdx.x] = v[q];
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if ((int)threadIdx.x < st) s[threadIdx.x] += s[threadIdx.x + st];
      __syncthreads();
    }
    if (threadIdx.x == 0) a.diag_partial[(size_t)blockIdx.x * KIDMP_NDIAG + q] = s[0];
    __syncthreads();
  }
}

// fixed-order reduction of the block partials into diag[8] (accumulates: kidmp_diag reads and clears)
__global__ void k_diag_reduce(const double* __restrict__ partial, int nblocks, double* __restrict__ diag) {
  __shared__ double s[256];
  const int q = blockIdx.x;
  double x = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) x += partial[(size_t)b * KIDMP_NDIAG + q];
  s[threadIdx.x] = x;
  __syncthreads();
  for (int st = blockDim.x >> 1; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) s[threadIdx.x] += s[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) diag[q] += s[0];
}

// layout conversion between KiD's (k,i) arrays [col][nz] and the device layout [nz][ncol]
__global__ void k_transpose(const float* __restrict__ src, float* __restrict__ dst, long ncol, int nz, int to_col_fastest) {
  __shared__ float tile[32][33];
  // src is [R][C] row-major, dst is [C][R]
  // blockIdx.x tiles the columns, blockIdx.y the levels
  const long R = to_col_fastest ? ncol : nz, Cn = to_col_fastest ? nz : ncol;
  const long tc = (long)blockIdx.x * 32, tk = (long)blockIdx.y * 32;
  const long r0 = to_col_fastest ? tc : tk, c0 = to_col_fastest ? tk : tc;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Cn) tile[j][threadIdx.x] = src[r * Cn + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Cn) dst[c * R + r] = tile[threadIdx.x][j];
  }
}

#undef R1
#undef R2
#undef EPSF
#undef T_0
#undef D0r
#undef D0c
#undef D0s
#undef D0g

}  // namespace kidmp
