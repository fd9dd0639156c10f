// mphys_thompson09n.hpp - C++ twin of KiD's interface module `mphys_thompson09n` (reference: I: =
// /root/reference/mphys_thompson09n.f90) over the C ABI of include/kidmp.h.
//
// The reference's host is Fortran and talks to the scheme through module variables, not arguments
// (I:11-17, I:28).  This image has no Fortran compiler, so the same host-side contract is restated in
// C++ with the same names: the KiD modules the interface `Use`s become namespaces holding the same
// variables, and `mphys_thompson09_interfacen()` takes no arguments, initialises the scheme on its
// first call (I:100-103), hands the nx columns to the GPU in one call and leaves the microphysics
// tendencies in `column_variables` and the surface precipitation diagnostics with `save_dg`
// (I:155-192, I:248-308).  kid_b200/fortran/mphys_thompson09n.f90 is the same thing in Fortran.
#pragma once
#include <functional>
#include <string>
#include <vector>

namespace kid {

// parameters.f90 (KiD host): array extents and names
namespace parameters {
extern int nx, nz;
extern float dt;
inline constexpr int nspecies = 5;                       // 1 cloud, 2 rain, 3 ice, 4 snow, 5 graupel (I:66-95)
extern std::string h_names[nspecies], mom_units[2];
}
namespace physconst { extern float p0, r_on_cp; }
namespace namelists { extern bool iiwarm; extern float set_Nc; }
namespace switches { extern bool l_sediment, l_reuse_thompson_lookup; }

// column_variables.f90 (KiD host).  Fortran (k,i) arrays are stored k fastest: index k + nz*i.
struct species { float moments[1][2]; };                  // %moments(bin, moment): 1 mass, 2 number
namespace column_variables {
extern std::vector<float> theta, qv, exner, dz;
extern std::vector<float> dtheta_adv, dtheta_div, dqv_adv, dqv_div, dtheta_mphys, dqv_mphys;
extern std::vector<species> hydrometeors, dhydrometeors_adv, dhydrometeors_div, dhydrometeors_mphys;   // (k,i,ih)
void allocate(int nx, int nz);
inline size_t ki(int k, int i) { return (size_t)k + (size_t)parameters::nz * i; }                      // 0-based
inline size_t kih(int k, int i, int ih) { return ki(k, i) + (size_t)parameters::nz * parameters::nx * ih; }
}

// diagnostics.f90 (KiD host): the hook the interface reports through.  value: nx numbers (1 for scalars).
namespace diagnostics {
using save_dg_fn = std::function<void(const std::vector<float>& value, const std::string& name, const std::string& units,
                                      const std::string& dim)>;
extern save_dg_fn save_dg;
extern int i_dgtime;
}

namespace mphys_thompson09n {
extern bool micro_unset;                                  // I:22
extern bool save_process_rates;                           // the 36 per-level save_dg rates of M:2963-3120 (default on)
extern int ndev;                                          // > 1 before the first call: the nx columns over several GPUs
// returns 0, or the non-zero status of the failing kidmp_* call (text: last_error())
int mphys_thompson09_interfacen();
inline int mphys_thompson09_interface() { return mphys_thompson09_interfacen(); }   // spelling used by BASELINE.json
const char* last_error();
void finalize();
}

}  // namespace kid
