// kidmp_kid.cuh - the KiD-facing algebra around the column step, on the device.
// Replaces the gather at I:59-97 (state + (advective + divergence tendency)*dt, theta -> T,
// exner -> p) and the scatter at I:198-245 (new state -> microphysics tendencies) of
// mphys_thompson09_interfacen.  KiD arrays are (k,i) = k fastest; the step's layout is column
// fastest, so both kernels transpose 32x32 tiles through shared memory: reads and writes are
// coalesced on both sides.
#pragma once
#include "kidmp_internal.h"
#include "kidmp_math.cuh"

namespace kidmp {

// planes of the hydrometeor moments in kidmp_kid_columns::hyd order -> state field index
// qc qr nr qi ni qs qg  ->  F_QC F_QR F_NR F_QI F_NI F_QS F_QG
__device__ __constant__ int c_hyd_field[7] = {1, 3, 7, 2, 6, 4, 5};

struct KidArgs {
  long nx; int nz;
  float dt, p0, ooroc;          // ooroc = 1./r_on_cp (I:61)
  int iiwarm;
  const float *theta, *dtheta_adv, *dtheta_div, *exner, *qv, *dqv_adv, *dqv_div;
  const float *hyd[7], *dhyd_adv[7], *dhyd_div[7];
  float *dtheta_mphys, *dqv_mphys, *dhyd_mphys[7];
  float* f[KIDMP_NFIELDS];      // state, [nz][nx]
  float* p;                     // [nz][nx]
};

// block (32, 8); blockIdx.x tiles the columns, blockIdx.y the levels
__global__ void k_kid_gather(KidArgs a) {
  __shared__ float tile[32][33];
  const long i0 = (long)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  auto emit = [&](float* dst, auto value) {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {        // j: column in tile, threadIdx.x: level
      const long i = i0 + j; const int k = k0 + threadIdx.x;
      if (i < a.nx && k < a.nz) tile[j][threadIdx.x] = value(i * a.nz + k);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {        // j: level in tile, threadIdx.x: column
      const long i = i0 + threadIdx.x; const int k = k0 + j;
      if (i < a.nx && k < a.nz) dst[(long)k * a.nx + i] = tile[threadIdx.x][j];
    }
    __syncthreads();
  };
  emit(a.f[8], [&](long o) { return (a.theta[o] + (a.dtheta_adv[o] + a.dtheta_div[o]) * a.dt) * a.exner[o]; });   // I:60
  emit(a.p, [&](long o) { return a.p0 * pow_f(a.exner[o], a.ooroc); });                                           // I:61
  emit(a.f[0], [&](long o) { return a.qv[o] + (a.dqv_adv[o] + a.dqv_div[o]) * a.dt; });                           // I:64
  for (int m = 0; m < 7; ++m) {
    const bool ice = m >= 3;
    const float *h = a.hyd[m], *ha = a.dhyd_adv[m], *hd = a.dhyd_div[m];
    if (ice && (a.iiwarm || !h)) emit(a.f[c_hyd_field[m]], [&](long) { return 0.0f; });                           // I:46-52, I:78
    else emit(a.f[c_hyd_field[m]], [&](long o) { return h[o] + (ha[o] + hd[o]) * a.dt; });                        // I:66-95
  }
}

__global__ void k_kid_scatter(KidArgs a) {
  __shared__ float tile[32][33];
  const long i0 = (long)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  auto emit = [&](const float* src, auto store) {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {        // j: level, threadIdx.x: column
      const long i = i0 + threadIdx.x; const int k = k0 + j;
      if (i < a.nx && k < a.nz) tile[threadIdx.x][j] = src[(long)k * a.nx + i];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {        // j: column, threadIdx.x: level
      const long i = i0 + j; const int k = k0 + threadIdx.x;
      if (i < a.nx && k < a.nz) store(i * a.nz + k, tile[j][threadIdx.x]);
    }
    __syncthreads();
  };
  emit(a.f[8], [&](long o, float t1d) {                                                                            // I:199-200
    a.dtheta_mphys[o] = (t1d / a.exner[o] - a.theta[o]) / a.dt - (a.dtheta_adv[o] + a.dtheta_div[o]);
  });
  emit(a.f[0], [&](long o, float q) { a.dqv_mphys[o] = (q - a.qv[o]) / a.dt - (a.dqv_adv[o] + a.dqv_div[o]); });   // I:205-206
  for (int m = 0; m < 7; ++m) {
    const float *h = a.hyd[m], *ha = a.dhyd_adv[m], *hd = a.dhyd_div[m];
    float* out = a.dhyd_mphys[m];
    if (m >= 3 && (a.iiwarm || !h || !out)) continue;                                                              // I:226
    emit(a.f[c_hyd_field[m]], [&](long o, float q) { out[o] = (q - h[o]) / a.dt - (ha[o] + hd[o]); });              // I:211-243
  }
}

}  // namespace kidmp
