"""CPU test of the table-driven f64 exp / log used by the column kernels (kid_b200/csrc/kidmp_fastmath.h):
the same header compiles for the host, so its accuracy is checked here against libm."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "kidmp_fastmath.h"
#include <cstdio>
#include <cmath>
#include <random>
using namespace kidmp;
int main() {
  static double et[KFM_N]; static LogNode lt[KFM_N];
  kfm_build_tables(et, lt);
  std::mt19937_64 g(1);
  std::uniform_real_distribution<double> ue(-700, 700), ul(-300, 300), us(-40, 40), ux(-30, 10), uy(-4, 8);
  double maxe = 0, maxl = 0, maxa = 0;
  for (long n = 0; n < 2000000; ++n) {
    double x = (n & 1) ? ue(g) : us(g);
    double a = kfm_exp(x, et), b = exp(x);
    maxe = fmax(maxe, fabs(a - b) / b);
    double y = (n % 3 == 0) ? 1.0 + us(g) * 1e-3 : exp(ul(g) * 2.302585);
    double c = kfm_log(y, lt), d = log(y);
    if (fabs(d) > 1e-3) maxl = fmax(maxl, fabs(c - d) / fabs(d)); else maxa = fmax(maxa, fabs(c - d));
  }
  long diff = 0;
  for (long n = 0; n < 1000000; ++n) {
    float x = (float)exp(ux(g) * 2.3), y = (float)uy(g);
    float a = (float)kfm_exp((double)y * kfm_log((double)x, lt), et), b = (float)exp((double)y * log((double)x));
    if (a != b) ++diff;
  }
  int special = (kfm_exp(-INFINITY, et) == 0.0) && std::isinf(kfm_exp(800.0, et)) && std::isinf(kfm_log(0.0, lt)) &&
                std::isnan(kfm_log(-1.0, lt)) && std::isnan(kfm_exp(NAN, et));
  printf("%.3e %.3e %.3e %ld %d\n", maxe, maxl, maxa, diff, special);
  return 0;
}
'''


def test_fast_exp_log_against_libm(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "kid_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    maxe, maxl, maxa, diff, special = float(out[0]), float(out[1]), float(out[2]), int(out[3]), int(out[4])
    assert maxe < 4e-16          # exp: within 4 ulp of f64
    assert maxl < 1e-15          # log: relative, away from 1
    assert maxa < 1e-17          # log: absolute, near 1
    assert diff <= 2             # f32 x**y through the fast pair rounds like the libm-f64 evaluation
    assert special == 1
