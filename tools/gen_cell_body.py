"""Generates kid_b200/csrc/kidmp_cell_body.inc, the carry-free cell code of the unit-parallel physics kernel, from the
level body of k_column_step in kidmp_column.cuh: S4, S10 and S13 lose their vertical carries (kidmp_units.cuh phases 1 and 3
take them over).  Run it after changing the cell code in kidmp_column.cuh; the two kernels must give the same bits
(tools/ab.sh, tests/test_gpu_parity.py::test_fused_and_split_steps_are_bit_identical)."""
import re
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, 'kid_b200', 'csrc', 'kidmp_column.cuh')).read()
L=src.split('\n')
def find(t,start=0):
    for i in range(start,len(L)):
        if t in L[i]: return i
    raise KeyError(t)
k1=find('__global__ void __launch_bounds__(WARPS * 32, MINB) k_column_step(StepArgs a) {')
b0=find('// rates, M:1184-1211 (zeroed M:1282-1363)',k1)
b1=find('        // hand-off to the sedimentation kernel: [SC_N][nz][ncol]',k1)
body='\n'.join(L[b0:b1])
def rep(old,new,cnt=1):
    global body
    assert body.count(old)==cnt,(body.count(old),old[:90])
    body=body.replace(old,new)
# S4
rep('''          if (temp >= 270.65f) warm_above_a = true;
          { double nm = N0_min_a; graupel_n0(!warm_above_a && k > 0, L_qr, L_qg, mvd_r, rg, n0_empty, nm, ilamg, N0_g); N0_min_a = (float)nm; }''',
'''          graupel_slope(n0_min_a, L_qg, rg, ilamg, N0_g);       // the running minimum of M:1648 comes from phase 1''')
# S10
rep('''          if (temp >= 270.65f) warm_above_b = true;
          { double nm = N0_min_b; graupel_n0(!warm_above_b && k > 0, L_qr, L_qg, mvd_r, rg, n0_empty, nm, ilamg, N0_g); N0_min_b = (float)nm; }''',
'''          // M:2721-2731: whether this level lies above k_0 depends on the updated temperatures of the levels above,
          // which phase 3 knows: both values of the intercept are handed to it (they differ only with supercooled rain),
          // and it evaluates the slope, which only the graupel fall speed of S13 reads
          warm9 = temp >= 270.65f;
          n0b_lo = (rg > 5.E-5f) ? graupel_n0_exp(0.01f, rg) : n0_empty;
          n0b_slw = (L_qr && mvd_r > 100.E-6f) ? graupel_n0_exp(4.01f + log10_f(mvd_r), rg) : n0b_lo;''')
# S13 rain
rep('''        } else {
          v_r = vtr_up; v_nr = vtnr_up;
        }
        if (fmaxf(v_r, v_nr) > 1.E-3f) {
          ksed_r = (unsigned short)max((int)ksed_r, k + 1);
          const float delta_tp = dzq / (fmaxf(v_r, v_nr));
          nstep_r = (unsigned short)min(max((int)nstep_r, (int)(DT / delta_tp + 1.f)), KP_NSTEP_MAX);
        }''','''        } else {
          v_r = 0.f; v_nr = 0.f;                              // phase 3: the speeds of the level above (M:3235)
        }''')
rep('''          } else {
            v_i = vti_up; v_ni = vtni_up;
          }
          if (v_i > 1.E-3f) {
            ksed_i = (unsigned short)max((int)ksed_i, k + 1);
            const float delta_tp = dzq / v_i;
            nstep_i = (unsigned short)min(max((int)nstep_i, (int)(DT / delta_tp + 1.f)), KP_NSTEP_MAX);
          }''','''          }''')
rep('''            if (temp > (T_0 + 0.1f)) v_s = fmaxf(vts * vts_boost, vts * ((v_r - vts * vts_boost) / (temp - T_0)));
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
''','''            vts_h = vts;                                      // M:3301 needs the rain speed after the rule of M:3235: phase 3
          }
        }
''')
rep("        float v_r, v_nr, v_i = 0.f, v_ni = 0.f, v_s = 0.f, v_g = 0.f;", "        float v_r, v_nr, v_i = 0.f, v_ni = 0.f;")
rep("        const float dzq = a.dz_col ? a.dz_col[o + col] : a.dz[k];\n", "")      # the layer depth is only read by phase 3

for bad in ('warm_above','vtr_up','vtnr_up','vti_up','vtni_up','vts_up','vtg_up','nstep_','ksed_','N0_min_'):
    assert bad not in body, bad
open(os.path.join(ROOT, 'kid_b200', 'csrc', 'kidmp_cell_body.inc'), 'w').write(body)
print(len(body.split('\n')),'body lines')
