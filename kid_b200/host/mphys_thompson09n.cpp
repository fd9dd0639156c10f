// mphys_thompson09n.cpp - see mphys_thompson09n.hpp.  Host code only: links against libkidmp.so.
#include "mphys_thompson09n.hpp"
#include "../../include/kidmp.h"

namespace kid {

namespace parameters {
int nx = 1, nz = 60;
float dt = 1.0f;
std::string h_names[nspecies] = {"cloud", "rain", "ice", "snow", "graupel"};
std::string mom_units[2] = {"kg/kg", "/kg"};
}
namespace physconst { float p0 = 100000.0f, r_on_cp = 287.05f / 1005.0f; }
namespace namelists { bool iiwarm = false; float set_Nc = 100.0f; }
namespace switches { bool l_sediment = true, l_reuse_thompson_lookup = false; }

namespace column_variables {
std::vector<float> theta, qv, exner, dz, dtheta_adv, dtheta_div, dqv_adv, dqv_div, dtheta_mphys, dqv_mphys;
std::vector<species> hydrometeors, dhydrometeors_adv, dhydrometeors_div, dhydrometeors_mphys;
void allocate(int nx_, int nz_) {
  parameters::nx = nx_; parameters::nz = nz_;
  const size_t n = (size_t)nx_ * nz_;
  for (auto* v : {&theta, &qv, &exner, &dtheta_adv, &dtheta_div, &dqv_adv, &dqv_div, &dtheta_mphys, &dqv_mphys}) v->assign(n, 0.f);
  dz.assign(nz_, 0.f);
  const species zero = {{{0.f, 0.f}}};
  for (auto* v : {&hydrometeors, &dhydrometeors_adv, &dhydrometeors_div, &dhydrometeors_mphys})
    v->assign(n * parameters::nspecies, zero);
}
}

namespace diagnostics {
save_dg_fn save_dg = nullptr;
int i_dgtime = 0;
}

namespace mphys_thompson09n {

bool micro_unset = true;
bool save_process_rates = true;
int ndev = 1;
static kidmp_handle* g_handle = nullptr;
static std::string g_error;
// planes handed to the library: (ih, imom) of hydrometeors(k,i,ih)%moments(1,imom), 0-based here
static const int plane_ih[7] = {0, 1, 1, 2, 2, 3, 4};
static const int plane_im[7] = {0, 0, 1, 0, 1, 0, 0};

const char* last_error() { return g_error.c_str(); }
void finalize() {
  if (g_handle) kidmp_finalize(g_handle);
  g_handle = nullptr; micro_unset = true;
}

int mphys_thompson09_interfacen() {
  using namespace column_variables;
  const int nx = parameters::nx, nz = parameters::nz;
  const size_t n = (size_t)nx * nz;
  // Initialise microphysics (I:100-103)
  if (micro_unset) {
    kidmp_config cfg{};
    cfg.set_Nc = namelists::set_Nc;
    cfg.iiwarm = namelists::iiwarm ? 1 : 0;
    cfg.l_sediment = switches::l_sediment ? 1 : 0;
    cfg.wp_double = 0;
    cfg.device = 0;
    cfg.reuse_tables = switches::l_reuse_thompson_lookup ? 1 : 0;
    cfg.table_cache_path = "run_data/kidmp_tables.bin";
    cfg.ndev = ndev;
    cfg.device_ids = nullptr;
    const int rc = kidmp_init(&cfg, &g_handle);
    if (rc) { g_error = kidmp_last_error(nullptr); return rc; }
    if (save_process_rates && kidmp_enable_rates(g_handle, 1)) { g_error = kidmp_last_error(g_handle); return 1; }
    micro_unset = false;
  }
  const int np = namelists::iiwarm ? 3 : 7;
  static std::vector<float> hyd[7], adv[7], dvg[7], out[7], ppt;
  for (int m = 0; m < np; ++m) {
    hyd[m].resize(n); adv[m].resize(n); dvg[m].resize(n); out[m].resize(n);
    for (int i = 0; i < nx; ++i)
      for (int k = 0; k < nz; ++k) {
        const size_t s = kih(k, i, plane_ih[m]), d = ki(k, i);
        hyd[m][d] = hydrometeors[s].moments[0][plane_im[m]];
        adv[m][d] = dhydrometeors_adv[s].moments[0][plane_im[m]];
        dvg[m][d] = dhydrometeors_div[s].moments[0][plane_im[m]];
      }
  }
  ppt.assign((size_t)nx * 4, 0.f);
  kidmp_kid_columns c{};
  c.nx = nx; c.nz = nz;
  c.theta = theta.data(); c.dtheta_adv = dtheta_adv.data(); c.dtheta_div = dtheta_div.data(); c.exner = exner.data();
  c.qv = qv.data(); c.dqv_adv = dqv_adv.data(); c.dqv_div = dqv_div.data(); c.dz = dz.data();
  for (int m = 0; m < 7; ++m) {
    c.hyd[m] = m < np ? hyd[m].data() : nullptr; c.dhyd_adv[m] = m < np ? adv[m].data() : nullptr;
    c.dhyd_div[m] = m < np ? dvg[m].data() : nullptr; c.dhyd_mphys[m] = m < np ? out[m].data() : nullptr;
  }
  c.dtheta_mphys = dtheta_mphys.data(); c.dqv_mphys = dqv_mphys.data(); c.ppt = ppt.data();
  // gather, mp_thompson for every column, back out tendencies (I:54-246): one call, on the GPU
  const int rc = kidmp_kid_interface(g_handle, &c, parameters::dt, physconst::p0, physconst::r_on_cp);
  if (rc) { g_error = kidmp_last_error(g_handle); return rc; }
  for (int m = 0; m < np; ++m)
    for (int i = 0; i < nx; ++i)
      for (int k = 0; k < nz; ++k)
        dhydrometeors_mphys[kih(k, i, plane_ih[m])].moments[0][plane_im[m]] = out[m][ki(k, i)];

  // the per-level process rates that mp_thompson itself saves (M:2963-3120): 36 save_dg calls per level in the reference's
  // order - column by column, level by level, the 30 ice-phase rates only when not iiwarm
  if (diagnostics::save_dg && save_process_rates) {
    static std::vector<float> rates;
    rates.resize(n * KIDMP_NRATES);
    if (kidmp_get_rates(g_handle, KIDMP_K_FASTEST, rates.data())) { g_error = kidmp_last_error(g_handle); return 1; }
    static std::vector<std::string> names;
    if (names.empty()) {
      std::string all = kidmp_rate_names();
      for (size_t a = 0; a < all.size();) { const size_t b = all.find(',', a); names.push_back(all.substr(a, b - a)); a = b == std::string::npos ? all.size() : b + 1; }
    }
    for (int i = 0; i < nx; ++i)
      for (int k = 0; k < nz; ++k)
        for (int r = namelists::iiwarm ? 30 : 0; r < KIDMP_NRATES; ++r)
          diagnostics::save_dg(std::vector<float>(1, rates[(size_t)r * n + ki(k, i)]), names[r], "/kg/s", nx == 1 ? "z" : "z,x");
  }

  // diagnostics as the reference saves them (I:155-192 for nx == 1, I:248-308 otherwise); ppt = rain, ice, snow, graupel
  if (diagnostics::save_dg) {
    const std::string units = parameters::mom_units[0] + " m";
    const int ih_of[4] = {1, 2, 3, 4};                                   // h_names index of rain, ice, snow, graupel
    std::vector<float> total(nx, 0.f);
    for (int q = 0; q < 4; ++q) {
      std::vector<float> v(ppt.begin() + (size_t)q * nx, ppt.begin() + (size_t)(q + 1) * nx);
      for (int i = 0; i < nx; ++i) total[i] += v[i];
      const std::string name = "surface_ppt_for_" + parameters::h_names[ih_of[q]];
      if (nx == 1) diagnostics::save_dg(v, name, units, "time");
      else {
        std::vector<float> mean(v);
        for (auto& x : mean) x /= (float)nx;
        diagnostics::save_dg(mean, name, units, "time");                 // I:255 ff.: column means ...
        diagnostics::save_dg(v, name, units, "time");                    // I:281 ff.: ... and every column
      }
    }
    if (nx == 1) diagnostics::save_dg(total, "total_surface_ppt", units, "time");
    else {
      std::vector<float> mean(total);
      for (auto& x : mean) x /= (float)nx;
      diagnostics::save_dg(mean, "total_surface_ppt", units, "time");
      diagnostics::save_dg(total, "total_surface_ppt", units, "time");
      diagnostics::save_dg(std::vector<float>((size_t)nx * nz, 0.f), "total_ppt_level", units, "z,x");   // never assigned in the reference (U3)
    }
  }
  return 0;
}

}  // namespace mphys_thompson09n
}  // namespace kid
