// kidmp_aero.cuh - the functions of the aerosol-aware half of the scheme (is_aerosol_aware = .true., M:28; dustyIce =
// homogIce = .true., M:30-31): Eff_aero M:4354-4390, iceDeMott M:4720-4756, iceKoop M:4764-4789, activ_ncloud M:4451-4526.
// f32 like the reference; transcendentals through the f64 routines of kidmp_math.cuh (rule 2).
#pragma once
#include "kidmp_math.cuh"

namespace kidmp {

#define KP_NAIN1 0.5E6f            /* naIN1, M:59 */

__device__ __forceinline__ float log_f(float x) { return (float)dlog((double)x); }

// collision efficiency of rain ('r'), snow ('s'), graupel ('g') with aerosols of diameter Da (Wang et al. 2010 after Slinn 1983)
template <char SPECIES>
__device__ __noinline__ float eff_aero(float D, float Da, float visc, float rhoa, float Temp) {
  const float boltzman = 1.3806503E-23f, meanPath = 0.0256E-6f;
  float vt = 1.f;
  if (SPECIES == 'r') vt = -0.1021f + 4.932E3f * D - 0.9551E6f * D * D + 0.07934E9f * D * D * D - 0.002362E12f * D * D * D * D;
  else if (SPECIES == 's') vt = KP_AV_S * pow_f(D, KP_BV_S);
  else if (SPECIES == 'g') vt = KP_AV_G * pow_f(D, KP_BV_G);
  const float Cc = 1.f + 2.f * meanPath / Da * (1.257f + 0.4f * exp_f(-0.55f * Da / meanPath));
  const float diff = boltzman * Temp * Cc / (3.f * KP_PI * visc * Da);
  const float Re = 0.5f * rhoa * D * vt / visc;
  const float Sc = visc / (rhoa * diff);
  const float St = Da * Da * vt * 1000.f / (9.f * visc * D);
  const float aval = 1.f + log_f(1.f + Re);
  const float St2 = (1.2f + 1.f / 12.f * aval) / (1.f + aval);
  float Eff = 4.f / (Re * Sc) * (1.f + 0.4f * sqrtf(Re) * pow_f(Sc, 0.3333f) + 0.16f * sqrtf(Re) * sqrtf(Sc))
              + 4.f * Da / D * (0.02f + Da / D * (1.f + 2.f * sqrtf(Re)));
  if (St > St2) Eff = Eff + pow_f((St - St2) / (St - St2 + 0.666667f), 1.5f);
  return fmaxf(1.E-5f, fminf(Eff, 1.0f));
}

// ice nuclei from dust, DeMott et al. (2010); the qv / qvs / qvsi arguments of the reference are unused
__device__ __forceinline__ float ice_demott(float tempc, float rho, float nifa) {
  const float rho_not0 = 101325.f / (287.05f * 273.15f);
  const float nifa_cc = nifa * rho_not0 * 1.E-6f / rho;
  float xni = (5.94e-5f * pow_f(-tempc, 3.33f)) * pow_f(nifa_cc, (-0.0264f * tempc) + 0.0033f);
  xni = xni * rho / rho_not0 * 1000.f;
  return fmaxf(0.f, xni);
}

// homogeneous freezing of deliquesced aerosols after Koop et al. (2001)
__device__ __forceinline__ float ice_koop(float temp, float qv, float qvs, float naero, float dt) {
  const float R_uni = 8.314f;
  const float ar_volume = 4.f / 3.f * KP_PI * (float)((double)2.5e-6f * (double)2.5e-6f * (double)2.5e-6f);   // (2.5e-6)**3 of a PARAMETER: one rounding
  float xni = 0.0f;
  const float satw = qv / qvs;
  const float mu_diff = 210368.0f + (131.438f * temp) - (3.32373E6f / temp) - (41729.1f * log_f(temp));
  const float a_w_i = exp_f(mu_diff / (R_uni * temp));
  const float delta_aw = satw - a_w_i;
  float log_J_rate = -906.7f + (8502.0f * delta_aw) - (26924.0f * delta_aw * delta_aw) + (29180.0f * delta_aw * delta_aw * delta_aw);
  log_J_rate = fminf(20.0f, log_J_rate);
  const float J_rate = pow10_f(log_J_rate);
  const float prob_h = fminf(1.f - exp_f(-J_rate * ar_volume * dt), 1.f);
  if (prob_h > 0.f) xni = fminf(prob_h * naero, 1000.E3f);
  return fmaxf(0.0f, xni);
}

// activated fraction of the CCN; tnccn_act is all ones in this reference (M:752-762), the interpolation arithmetic is kept
__device__ __noinline__ float activ_ncloud(float Tt, float Ww, float NCCN) {
  const float ta_Na[8] = {0.f, 10.0f, 31.6f, 100.0f, 316.0f, 1000.0f, 3160.0f, 10000.0f};
  const float ta_Ww[10] = {0.f, 0.01f, 0.0316f, 0.1f, 0.316f, 1.0f, 3.16f, 10.0f, 31.6f, 100.0f};
  const int ntb_arc = 7, ntb_arw = 9;
  float n_local = NCCN * 1.E-6f, w_local = Ww;
  if (n_local >= ta_Na[ntb_arc]) n_local = ta_Na[ntb_arc] - 1.0f;
  else if (n_local <= ta_Na[1]) n_local = ta_Na[1] + 1.0f;
  int n;
  for (n = 2; n <= ntb_arc; ++n) if (n_local >= ta_Na[n - 1] && n_local < ta_Na[n]) break;
  const int i = min(n, ntb_arc);                       // (a NaN falls through the search: the reference then reads past the table)
  const float x1 = log_f(ta_Na[i - 1]), x2 = log_f(ta_Na[i]);
  if (w_local >= ta_Ww[ntb_arw]) w_local = ta_Ww[ntb_arw] - 1.0f;
  else if (w_local <= ta_Ww[1]) w_local = ta_Ww[1] + 0.001f;
  for (n = 2; n <= ntb_arw; ++n) if (w_local >= ta_Ww[n - 1] && w_local < ta_Ww[n]) break;
  const int j = min(n, ntb_arw);
  const float y1 = log_f(ta_Ww[j - 1]), y2 = log_f(ta_Ww[j]);
  (void)Tt;
  const float A = 1.0f, B = 1.0f, C = 1.0f, D = 1.0f;
  const float nx = log_f(n_local), wy = log_f(w_local);
  const float t = (nx - x1) / (x2 - x1), u = (wy - y1) / (y2 - y1);
  const float fraction = (1.0f - t) * (1.0f - u) * A + t * (1.0f - u) * B + t * u * C + (1.0f - t) * u * D;
  return NCCN * fraction;
}

}  // namespace kidmp
