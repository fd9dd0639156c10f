// kidmp_math.cuh - arithmetic helpers of the column kernels.
//
// Parity contract (DESIGN.md "Arithmetic"): the reference (M: = module_mp_thompson09n.f90) mixes
// default REAL (f32) and DOUBLE PRECISION (f64) and the result of many expressions steers a
// discontinuous table index or threshold.  The kernels therefore
//   * are compiled with -fmad=false, IEEE division and square root, no flush-to-zero, so every
//     + - * / sqrt rounds exactly like the Fortran expression it restates;
//   * evaluate every transcendental whose result is stored as f32 in f64 and round once, which
//     is within one f32 ulp of any faithful libm (gfortran calls glibc powf/expf/logf/log10f);
//   * keep f64 where the reference has DOUBLE PRECISION, with x**y as exp(y*log(x)) (relative
//     error ~1e-15, invisible once the rates are narrowed to the f32 tendencies) and the
//     PARAMETER exponents of the scheme (mu_r = mu_g = mu_i = 0, bm_r = bm_g = bm_i = 3,
//     bv_r = bv_i = 1, M:56-111) strength-reduced to products and square roots.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "kidmp_fastmath.h"

namespace kidmp {

// node tables of kidmp_fastmath.h, filled by kidmp_init (read-only afterwards, L1-resident)
__device__ double g_exp_tab[KFM_N];
__device__ LogNode g_log_tab[KFM_N];

// One copy of the f64 logarithm and exponential in the kernel image: the column kernel has ~130
// call sites and its working set must stay near the instruction cache (profiles/r01: inlining
// them made a 254 KB kernel that stalled 13 cycles per issue on instruction fetch).
__device__ __noinline__ double dlog(double x) { return kfm_log(x, g_log_tab); }
__device__ __noinline__ double dexp(double x) { return kfm_exp(x, g_exp_tab); }
// x**y for x > 0 (x = 0 gives 0 for y > 0, NaN propagates), f64
__device__ __forceinline__ double pow_d(double x, double y) { return dexp(y * dlog(x)); }
// f32 result: REAL ** REAL of the reference (a libm powf call under gfortran)
__device__ __forceinline__ float pow_f(float x, float y) { return (float)dexp((double)y * dlog((double)x)); }
__device__ __forceinline__ float exp_f(float x) { return (float)dexp((double)x); }
__device__ __forceinline__ float log10_f(float x) { return (float)(dlog((double)x) * 0.43429448190325182765); }
// 10.**y with REAL y (M:1560, M:1646, M:2242 ...)
__device__ __forceinline__ float pow10_f(float y) { return (float)dexp((double)y * 2.302585092994046); }
__device__ __forceinline__ float cube_f(float x) { double d = x; return (float)(d * d * d); }   // x**bm_r etc.
__device__ __forceinline__ double cube_d(double x) { return x * x * x; }
__device__ __forceinline__ double sq_d(double x) { return x * x; }

__device__ __forceinline__ int nint_f(float x) { return (int)lroundf(x); }     // NINT: half away from zero
__device__ __forceinline__ int nint_d(double x) { return (int)lround(x); }

// RSLF / RSIF, M:4656-4717 (Flatau et al. polynomials, Horner form, f32)
__device__ __noinline__ float rslf(float P, float T) {
  const float C0 = .611583699E03f, C1 = .444606896E02f, C2 = .143177157E01f, C3 = .264224321E-1f,
              C4 = .299291081E-3f, C5 = .203154182E-5f, C6 = .702620698E-8f, C7 = .379534310E-11f,
              C8 = -.321582393E-13f;
  float X = fmaxf(-80.f, T - 273.16f);
  float ESL = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
  ESL = fminf(ESL, P * 0.15f);
  return .622f * ESL / (P - ESL);
}
__device__ __noinline__ float rsif(float P, float T) {
  const float C0 = .609868993E03f, C1 = .499320233E02f, C2 = .184672631E01f, C3 = .402737184E-1f,
              C4 = .565392987E-3f, C5 = .521693933E-5f, C6 = .307839583E-7f, C7 = .105785160E-9f,
              C8 = .161444444E-12f;
  float X = fmaxf(-80.f, T - 273.16f);
  float ESI = C0 + X * (C1 + X * (C2 + X * (C3 + X * (C4 + X * (C5 + X * (C6 + X * (C7 + X * C8)))))));
  ESI = fminf(ESI, P * 0.15f);
  return .622f * ESI / (P - ESI);
}

}  // namespace kidmp
