// TEST INFRASTRUCTURE -- not part of the product.
//
// The UNMODIFIED reference `fix eph/coloured/exp` (/root/reference/fix_eph_coloured_exp.cpp: `fix eph` model 4 with an
// exponential memory kernel on both forces), reached by include path, never copied, behind the same C driver as the
// reference `fix eph` (oracle/ref/ref_driver.cpp).  Built into oracle/_ref/libeph_coloured_ref.so; pins the coloured
// branch of the restatement (oracle/eph_oracle.c, orc_fix_set_colour) and generates tests/golden/coloured_case.npz.
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iostream>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include "lammps_shim.h"

#define private public
#define protected public
#include "eph_spline.h"
#include "eph_linear.h"
#include "eph_beta.h"
#include "eph_fdm.h"
#include "fix_eph_coloured_exp.h"
#undef private
#undef protected

#include "fix_driver.h"

namespace {

struct RefColouredFix : LAMMPS_NS::FixEPHColouredExp {
  using LAMMPS_NS::FixEPHColouredExp::FixEPHColouredExp;
  // which: 0 rho[nt] 1 w 2 xi 3 f_EPH 4 f_RNG 5 f_dis 6 f_sto  (all but 0: [nl][3])
  void probe_copy(int which, size_t nl, size_t nt, double *out) {
    switch (which) {
      case 0: std::copy(rho_i, rho_i + nt, out); break;
      case 1: std::copy(&w_i[0][0], &w_i[0][0] + 3 * nl, out); break;
      case 2: std::copy(&xi_i[0][0], &xi_i[0][0] + 3 * nl, out); break;
      case 3: std::copy(&f_EPH[0][0], &f_EPH[0][0] + 3 * nl, out); break;
      case 4: std::copy(&f_RNG[0][0], &f_RNG[0][0] + 3 * nl, out); break;
      case 5: std::copy(&f_dis_i[0][0], &f_dis_i[0][0] + 3 * nl, out); break;
      case 6: std::copy(&f_sto_i[0][0], &f_sto_i[0][0] + 3 * nl, out); break;
      default: throw std::runtime_error("bad probe id");
    }
  }
  size_t grid_size() { return fdm.ntotal; }
  void grid_T(double *out) { std::copy(fdm.T_e.begin(), fdm.T_e.end(), out); }
};

}  // namespace

SHIM_DRIVER_DEFINE(refc, RefColouredFix)
