// kidmp_units.cuh - the column physics S1..S13 with (32 neighbouring cloudy columns) x (ONE level) as a warp's unit of work.
//
// Why: the cell code is bound by the length of its dependent instruction chains (profiles/r01_ncu_step_kernels.md): a
// block needs ~30 us per level whatever it holds, so walking the 60 levels of a column one after the other costs 1.8 ms
// per block however many of them are empty, and small domains (KiD's own cases) leave most warps of the GPU without work.
// Here a block still owns up to WARPS groups of 32 cloudy columns, but
//   phase 1  walks its columns top-down once with a few instructions per level: which (group, level) units hold a busy
//            cell (a hydrometeor or supersaturation: otherwise every process rate is zero, M:1676-2286, M:2780, M:2880), and
//            the running minimum of the graupel intercept of S4 (M:1639-1648), which only needs the level's inputs;
//   phase 2  runs the cell code on the busy units only, level-major, WARPS units per round, with no vertical carry at all;
//   phase 3  walks the columns top-down again and settles what does run down the column: the graupel intercept minimum of
//            S10 (M:2721-2731; only the graupel fall speed reads it), the fall speeds of levels without the species
//            (M:3235, M:3267, M:3307, M:3333), snow above 0 C and graupel, which need the rain speed after that rule (M:3301,
//            M:3328), the sub-step counts and top sedimenting levels (M:3242, M:3208), and the hand-off of the units that
//            phase 2 skipped (tendencies 0, contents R1 / R2, speeds from above).
// The lanes of a warp stay neighbouring columns of one level (coalesced, same table entries, same branches) and the
// hand-off to k_sediment is what k_column_step writes, bit for bit.
#pragma once
#include "kidmp_column.cuh"

namespace kidmp {

// M:1649-1653 for a given intercept (the running minimum already taken)
__device__ __forceinline__ void graupel_slope(double N0_exp, bool L_qg, float rg, double& ilamg, double& N0_g) {
  if (L_qg) {
    const double lam_exp = sqrt(sqrt(N0_exp * (double)ck.am_g * (double)ck.cgg[0] / (double)rg));   // **oge1, oge1 = 1/4
    const double lamg = lam_exp * (double)ck.lamg_fac;
    ilamg = (double)1.f / lamg;
    N0_g = N0_exp / ((double)ck.cgg[1] * lam_exp) * lamg;
  }
}

constexpr int KU_MAXNZ = 256;
__host__ __device__ constexpr int ku_smem_bytes(int threads, int nz) {
  return threads * 9 * 4 + (threads / 32) * nz * 4 + (threads / 32) * nz * 2 + 64;
}

#define R1 KP_R1
#define R2 KP_R2
#define EPSF KP_EPS
#define T_0 KP_T_0
#define D0r KP_D0R
#define D0c KP_D0C
#define D0s KP_D0S
#define D0g KP_D0G

// PACK: the busy cells of a level are packed, in column order, into as few warps as they need (the lanes of a warp are
// then the busy cells of neighbouring columns, with the empty ones between them left out); without it a unit is a whole
// group of 32 columns as soon as one of its cells is busy.
template <int WARPS, int MINB, int BARS, bool RATES, bool PACK>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_unit_step(StepArgs a) {
  constexpr int NT = WARPS * 32;
  const int count = *a.work_count;                     // cloudy columns, compacted: every group but the last is full
  // the groups of the list are dealt evenly to a whole number of waves of blocks (see k_column_step)
  const int total_warps = (count + 31) >> 5;
  int nblocks = (total_warps + WARPS - 1) / WARPS;
  const int wave = a.nsm * MINB;
  nblocks = min((int)gridDim.x, (nblocks + wave - 1) / wave * wave);
  if ((int)blockIdx.x >= nblocks) return;
  const int w0 = (int)((long)blockIdx.x * total_warps / nblocks), w1 = (int)((long)(blockIdx.x + 1) * total_warps / nblocks);
  const int ng = w1 - w0;                              // groups of this block, <= WARPS
  if (ng <= 0) return;
  const int nz = a.nz;
  const long ncol = a.ncol;
  const float DT = a.dt;
  const float odt = 1.f / DT, odts = 1.f / DT;
  const float Nt_c = ck.Nt_c;
  const bool iiwarm = ck.iiwarm != 0;
  constexpr bool LOCKSTEP = WARPS > 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long ss = (long)nz * ncol;
  extern __shared__ float smem_ku[];
  float* const s_in = smem_ku;                                                 // [9][NT] inputs of the cell, parked over S3..S7
  unsigned* const s_cmask = reinterpret_cast<unsigned*>(s_in + 9 * NT);       // [WARPS][nz] busy lanes of every (group, level)
  unsigned short* const s_unit = reinterpret_cast<unsigned short*>(s_cmask + WARPS * nz);   // [<= WARPS*nz] units, level-major: k * WARPS + (group | packed warp)
  __shared__ int s_nunits;
  for (int i = tid; i < WARPS * nz; i += NT) s_cmask[i] = 0u;
  const double n0_empty = graupel_n0_lo();
  __syncthreads();

  // ================= phase 1: which units are busy; graupel intercept minimum of S4 ==========================
  if (warp < ng) {
    const int slot = (w0 + warp) * 32 + lane;
    const bool active = slot < count;
    const long col = (long)a.work_list[active ? slot : (w0 + warp) * 32];
    bool warm_a = false;                               // a level at or above this one has T >= 270.65 K (k_0, M:1635)
    double n0_min = (double)KP_GONV_MAX;
#pragma unroll 1
    for (int k = nz - 1; k >= 0; --k) {
      const long g = (long)k * ncol + col;
      const float t1d = a.f[F_T][g], pres = a.p[g], qv = fmaxf(1.E-10f, a.f[F_QV][g]);
      const float qc1d = a.f[F_QC][g], qi1d = a.f[F_QI][g], qr1d = a.f[F_QR][g], qs1d = a.f[F_QS][g], qg1d = a.f[F_QG][g];
      bool busy = qc1d > R1 || qi1d > R1 || qr1d > R1 || qs1d > R1 || qg1d > R1;
      if (!busy) {
        const float tempc = t1d - 273.15f;
        const float qvs = rslf(pres, t1d);
        const float qvsi = (tempc <= 0.0f) ? rsif(pres, t1d) : qvs;
        float ssatw = qv / qvs - 1.f;
        float ssati = qv / qvsi - 1.f;
        if (fabsf(ssatw) < EPSF) ssatw = 0.0f;
        if (fabsf(ssati) < EPSF) ssati = 0.0f;
        busy = ssati > 0.0f || ssatw > EPSF;
      }
      if (!iiwarm) {
        if (t1d >= 270.65f) warm_a = true;
        const float rho = 0.622f * pres / (KP_R * t1d * (qv + 0.622f));
        const float rg = (qg1d > R1) ? qg1d * rho : R1;
        bool slw = false;
        float mvd_r = 0.f;
        if (!warm_a && k > 0 && qr1d > R1) {             // the rain of S1 (M:1445-1466) for the xslw1 of M:1640
          const float rr = qr1d * rho;
          float nr = fmaxf(R2, a.f[F_NR][g] * rho);
          if (nr <= R2) { mvd_r = 1.0E-3f; nr = nr_from_mvd(rr, mvd_r); }
          const double lamr = rain_lam(nr, rr);
          mvd_r = (float)((double)(3.0f + 0.f + 0.672f) / lamr);
          if (mvd_r > 2.5E-3f) mvd_r = 2.5E-3f;
          else if (mvd_r < D0r * 0.75f) mvd_r = D0r * 0.75f;
          slw = mvd_r > 100.E-6f;
        }
        double N0_exp = n0_empty;
        if (slw || rg > 5.E-5f) N0_exp = graupel_n0_exp(slw ? 4.01f + log10_f(mvd_r) : 0.01f, rg);
        n0_min = fmin(N0_exp, n0_min);
        if (active) a.scratch[SC_N0A * ss + g] = (float)n0_min;     // values of M:1646 are f32 numbers: exact
      }
      const unsigned bm = __ballot_sync(0xffffffffu, busy && active);
      if (lane == 0) s_cmask[warp * nz + k] = bm;
    }
  }
  __syncthreads();
  // ---- the units of the block, level-major from the top: a round of phase 2 holds units of one or two levels ----
  if (warp == 0) {
    int n = 0;
    for (int k = nz - 1; k >= 0; --k) {
      const unsigned cm = lane < ng ? s_cmask[lane * nz + k] : 0u;
      if (PACK) {
        int cells = __popc(cm);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, d);
        const int nw_k = (cells + 31) >> 5;              // <= WARPS
        if (lane < nw_k) s_unit[n + lane] = (unsigned short)(k * WARPS + lane);
        n += nw_k;
      } else {
        const unsigned m = __ballot_sync(0xffffffffu, cm != 0u);
        if (cm != 0u) s_unit[n + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(k * WARPS + lane);
        n += __popc(m);
      }
    }
    if (lane == 0) s_nunits = n;
  }
  __syncthreads();
  const int nunits = s_nunits;

  // ================= phase 2: S1..S13 on the busy units, WARPS at a time =========================================
#define LOCKBAR(i) do { if (LOCKSTEP && ((BARS >> (i)) & 1)) { __syncwarp(); asm volatile("bar.sync 1, %0;" ::"r"(lock_threads) : "memory"); } } while (0)
#pragma unroll 1
  for (int base = 0; base < nunits; base += WARPS) {
    if (base > 0) __syncthreads();                     // the stage barriers of two rounds must not mix
    const int busy_warps = min(WARPS, nunits - base);
    if (warp >= busy_warps) continue;
    const int lock_threads = busy_warps * 32;
    const int unit = s_unit[base + warp];
    const int k = unit / WARPS, gi = unit - k * WARPS;
    int slot;
    bool active;
    if (PACK) {
      // lane -> the (gi*32 + lane)-th busy cell of level k among the block's columns, in column order; lanes past the
      // last busy cell shadow the warp's first one: same branches, nothing stored
      int want = gi * 32 + lane, g = 0, before = 0, g_first = -1, r_first = 0;
      int found_g = -1, found_r = 0;
      for (g = 0; g < ng; ++g) {
        const int c = __popc(s_cmask[g * nz + k]);
        if (g_first < 0 && gi * 32 < before + c) { g_first = g; r_first = gi * 32 - before; }
        if (found_g < 0 && want < before + c) { found_g = g; found_r = want - before; }
        before += c;
      }
      active = found_g >= 0;
      if (!active) { found_g = g_first; found_r = r_first; }
      slot = (w0 + found_g) * 32 + (int)__fns(s_cmask[found_g * nz + k], 0u, found_r + 1);
    } else {
      slot = (w0 + gi) * 32 + lane;
      active = slot < count;
      // lanes past the end of the list shadow their group's first column: same branches, nothing stored
      if (!active) slot = (w0 + gi) * 32;
    }
    const long col = (long)a.work_list[slot];
    LOCKBAR(0);
    {
      {
        const long o = (long)k * ncol;
        const float t1d = a.f[F_T][o + col], qv1d = a.f[F_QV][o + col], pres = a.p[o + col];
        float qc1d = a.f[F_QC][o + col], qi1d = a.f[F_QI][o + col], qr1d = a.f[F_QR][o + col], qs1d = a.f[F_QS][o + col],
              qg1d = a.f[F_QG][o + col];
        float ni1d = a.f[F_NI][o + col], nr1d = a.f[F_NR][o + col];
        const double n0_min_a = iiwarm ? (double)KP_GONV_MAX : (double)a.scratch[SC_N0A * ss + o + col];
        bool warm9 = false;
        double n0b_lo = n0_empty, n0b_slw = n0_empty;
        float vts_h = 0.f;
#include "kidmp_cell_body.inc"
        // hand-off: what k_column_step writes, except that the fall speeds are this level's own (0 without the species:
        // phase 3 fills in the rule of the level above, snow above 0 C and graupel) plus five values for phase 3
        if (active) {
          float s15 = 0.0f;
          if (temp > T_0) s15 = ck.lfus * ocp;
          else if (temp < KP_HGFR) s15 = -((KP_LSUB - lvap) * ocp);
          float* sc = a.scratch + o + col;
          sc[SC_TTEN * ss] = tt; sc[SC_QVTEN * ss] = qvt; sc[SC_QCTEN * ss] = qct; sc[SC_QITEN * ss] = qit;
          sc[SC_QRTEN * ss] = qrt; sc[SC_QSTEN * ss] = qst; sc[SC_QGTEN * ss] = qgt; sc[SC_NITEN * ss] = nit;
          sc[SC_NRTEN * ss] = nrt; sc[SC_NCTEN * ss] = nct;
          sc[SC_RR * ss] = rr; sc[SC_NR * ss] = nr; sc[SC_RI * ss] = ri; sc[SC_NI * ss] = ni; sc[SC_RS * ss] = rs; sc[SC_RG * ss] = rg;
          sc[SC_VTR * ss] = v_r; sc[SC_VTNR * ss] = v_nr; sc[SC_VTI * ss] = v_i; sc[SC_VTNI * ss] = v_ni;
          sc[SC_RHO * ss] = rho; sc[SC_S15 * ss] = s15;
          // the sign of the first intercept carries the k_0 test of this level's updated temperature (intercepts are > 0)
          sc[SC_N0A * ss] = warm9 ? -(float)n0b_lo : (float)n0b_lo; sc[SC_N0B_SLW * ss] = (float)n0b_slw;
          sc[SC_VTS_RAW * ss] = vts_h; sc[SC_VTS_BOOST * ss] = vts_boost; sc[SC_TEMP * ss] = temp;
        }
        }   // shadowed inputs
      }
    }
  }
#undef LOCKBAR
  __syncthreads();

  // ================= phase 3: what runs down the column ==========================================================
  if (warp < ng) {
    const int slot = (w0 + warp) * 32 + lane;
    const bool active = slot < count;
    const long col = (long)a.work_list[active ? slot : (w0 + warp) * 32];
    double n0_min = (double)KP_GONV_MAX;               // M:2731
    bool warm_b = false;
    float v_r = 0.f, v_nr = 0.f, v_i = 0.f, v_ni = 0.f, v_s = 0.f, v_g = 0.f;     // speeds of the level above
    int nstep_r = 0, nstep_i = 0, nstep_s = 0, nstep_g = 0, ksed_r = 1, ksed_i = 1, ksed_s = 1, ksed_g = 1;
#pragma unroll 1
    for (int k = nz - 1; k >= 0; --k) {
      const long o = (long)k * ncol;
      float* sc = a.scratch + o + col;
      const float dzq = a.dz_col ? a.dz_col[o + col] : a.dz[k];
      const unsigned cm = s_cmask[warp * nz + k];
      const bool handed = PACK ? ((cm >> lane) & 1u) != 0u : cm != 0u;
      float rho = 0.f, s15 = 0.f;
      if (handed) {
        const float rr = sc[SC_RR * ss], ri = sc[SC_RI * ss], rs = sc[SC_RS * ss], rg = sc[SC_RG * ss];
        if (rr > R1) { v_r = sc[SC_VTR * ss]; v_nr = sc[SC_VTNR * ss]; }
        if (!iiwarm) {
          const float x1 = sc[SC_N0A * ss];
          if (x1 < 0.f) warm_b = true;
          const double N0_exp = (!warm_b && k > 0) ? (double)sc[SC_N0B_SLW * ss] : (double)fabsf(x1);
          n0_min = fmin(N0_exp, n0_min);
          if (ri > R1) { v_i = sc[SC_VTI * ss]; v_ni = sc[SC_VTNI * ss]; }
          if (rs > R1) {
            const float vts = sc[SC_VTS_RAW * ss], vts_boost = sc[SC_VTS_BOOST * ss], temp = sc[SC_TEMP * ss];
            if (temp > (T_0 + 0.1f)) v_s = fmaxf(vts * vts_boost, vts * ((v_r - vts * vts_boost) / (temp - T_0)));
            else v_s = vts * vts_boost;
          }
          if (rg > R1) {
            rho = sc[SC_RHO * ss]; s15 = sc[SC_S15 * ss];
            const float rhof = sqrtf(ck.rho_not / rho);
            double ilamg = 0., N0_g = 0.;
            graupel_slope(n0_min, true, rg, ilamg, N0_g);
            const float vtg = (float)((double)(rhof * KP_AV_G * ck.cgg[5] * ck.ogg3) * pow_d(ilamg, (double)KP_BV_G));
            v_g = (s15 > 0.f) ? fmaxf(vtg, v_r) : vtg;         // temp > T_0 is what makes the S15 factor positive
          }
        }
      } else if (!iiwarm) {
        if (a.f[F_T][o + col] >= 270.65f) warm_b = true;
        n0_min = fmin(n0_empty, n0_min);
      }
      if (fmaxf(v_r, v_nr) > 1.E-3f) {
        ksed_r = max(ksed_r, k + 1);
        const float delta_tp = dzq / (fmaxf(v_r, v_nr));
        nstep_r = max(nstep_r, (int)(DT / delta_tp + 1.f));
      }
      if (!iiwarm) {
        if (v_i > 1.E-3f) { ksed_i = max(ksed_i, k + 1); const float d = dzq / v_i; nstep_i = max(nstep_i, (int)(DT / d + 1.f)); }
        if (v_s > 1.E-3f) { ksed_s = max(ksed_s, k + 1); const float d = dzq / v_s; nstep_s = max(nstep_s, (int)(DT / d + 1.f)); }
        if (v_g > 1.E-3f) { ksed_g = max(ksed_g, k + 1); const float d = dzq / v_g; nstep_g = max(nstep_g, (int)(DT / d + 1.f)); }
      }
      if (active) {
        if (!handed) {                                   // a unit without a busy cell: all rates zero, state unchanged
          const float temp = a.f[F_T][o + col], pres = a.p[o + col];
          const float qv = fmaxf(1.E-10f, a.f[F_QV][o + col]);
          const float tempc = temp - 273.15f;
          const float ocp = 1.f / (KP_CP * (1.f + 0.887f * qv));
          const float lvap = KP_LVAP0 + (2106.0f - 4218.0f) * tempc;
          s15 = 0.0f;
          if (temp > T_0) s15 = ck.lfus * ocp;
          else if (temp < KP_HGFR) s15 = -((KP_LSUB - lvap) * ocp);
#pragma unroll
          for (int q = SC_TTEN; q <= SC_NCTEN; ++q) sc[q * ss] = 0.0f;
          sc[SC_RR * ss] = R1; sc[SC_NR * ss] = R2; sc[SC_RI * ss] = R1; sc[SC_NI * ss] = R2; sc[SC_RS * ss] = R1; sc[SC_RG * ss] = R1;
          sc[SC_RHO * ss] = 0.622f * pres / (KP_R * temp * (qv + 0.622f)); sc[SC_S15 * ss] = s15;
          if (RATES && a.rates) {
            float* rp = a.rates + o + col;
            for (int q = 0; q < KIDMP_NRATES; ++q) rp[q * ss] = 0.0f;
          }
        }
        sc[SC_VTR * ss] = v_r; sc[SC_VTNR * ss] = v_nr; sc[SC_VTI * ss] = v_i; sc[SC_VTNI * ss] = v_ni;
        sc[SC_VTS * ss] = v_s; sc[SC_VTG * ss] = v_g;
      }
    }
    // column summary for the sedimentation kernel (U12: the sub-step count is capped, see k_column_step)
    if (active) {
      int* ci = a.colint + col;
      ci[0] = min(nstep_r, KP_NSTEP_MAX); ci[ncol] = min(nstep_i, KP_NSTEP_MAX); ci[2 * ncol] = min(nstep_s, KP_NSTEP_MAX);
      ci[3 * ncol] = min(nstep_g, KP_NSTEP_MAX);
      ci[4 * ncol] = ksed_r; ci[5 * ncol] = ksed_i; ci[6 * ncol] = ksed_s; ci[7 * ncol] = ksed_g;
      if (max(max(nstep_r, nstep_i), max(nstep_s, nstep_g)) > 1) atomicAdd(a.redo_count, 1);
    }
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
