// kid_driver.cpp - tiny stand-in for KiD's time loop, used by tests/test_gpu_parity.py::test_cxx_host_twin:
// reads a column state written by the test (raw float32), fills kid::column_variables, calls
// kid::mphys_thompson09n::mphys_thompson09_interfacen() like KiD's mphys dispatch would, and writes the
// tendencies and the save_dg records back for the test to compare with the library's Python binding.
#include <cstdio>
#include <cstdlib>
#include <map>
#include "mphys_thompson09n.hpp"
using namespace kid;

static std::vector<float> rd(FILE* f, size_t n) { std::vector<float> v(n); if (fread(v.data(), 4, n, f) != n) { perror("read"); exit(2); } return v; }
static void wr(FILE* f, const std::vector<float>& v) { fwrite(v.data(), 4, v.size(), f); }

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: kid_driver in out nx nz dt iiwarm\n"); return 2; }
  const int nx = atoi(argv[3]), nz = atoi(argv[4]);
  parameters::dt = (float)atof(argv[5]);
  namelists::iiwarm = atoi(argv[6]) != 0;
  namelists::set_Nc = namelists::iiwarm ? 50.0f : 100.0f;
  column_variables::allocate(nx, nz);
  using namespace column_variables;
  const size_t n = (size_t)nx * nz;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  theta = rd(f, n); dtheta_adv = rd(f, n); dtheta_div = rd(f, n); exner = rd(f, n); qv = rd(f, n); dqv_adv = rd(f, n); dqv_div = rd(f, n);
  dz = rd(f, nz);
  const int pih[7] = {0, 1, 1, 2, 2, 3, 4}, pim[7] = {0, 0, 1, 0, 1, 0, 0};
  for (int m = 0; m < 7; ++m) {
    std::vector<float> h = rd(f, n), a = rd(f, n), d = rd(f, n);
    for (int i = 0; i < nx; ++i) for (int k = 0; k < nz; ++k) {
      hydrometeors[kih(k, i, pih[m])].moments[0][pim[m]] = h[ki(k, i)];
      dhydrometeors_adv[kih(k, i, pih[m])].moments[0][pim[m]] = a[ki(k, i)];
      dhydrometeors_div[kih(k, i, pih[m])].moments[0][pim[m]] = d[ki(k, i)];
    }
  }
  fclose(f);
  std::map<std::string, std::vector<float>> dg;
  int ncalls = 0;
  diagnostics::save_dg = [&](const std::vector<float>& v, const std::string& name, const std::string&, const std::string&) {
    ++ncalls; dg[name + "#" + std::to_string(v.size())] = v;
  };
  const int rc = mphys_thompson09n::mphys_thompson09_interfacen();      // no arguments, like the reference (I:28)
  if (rc) { fprintf(stderr, "interface failed: %s\n", mphys_thompson09n::last_error()); return 1; }
  FILE* o = fopen(argv[2], "wb");
  wr(o, dtheta_mphys); wr(o, dqv_mphys);
  for (int m = 0; m < 7; ++m) {
    std::vector<float> t(n);
    for (int i = 0; i < nx; ++i) for (int k = 0; k < nz; ++k) t[ki(k, i)] = dhydrometeors_mphys[kih(k, i, pih[m])].moments[0][pim[m]];
    wr(o, t);
  }
  const char* names[4] = {"rain", "ice", "snow", "graupel"};
  for (int q = 0; q < 4; ++q) wr(o, dg["surface_ppt_for_" + std::string(names[q]) + "#" + std::to_string(nx)]);
  fclose(o);
  printf("ok save_dg_calls=%d micro_unset=%d\n", ncalls, (int)mphys_thompson09n::micro_unset);
  mphys_thompson09n::finalize();
  return 0;
}
