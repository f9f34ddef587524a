// TEST INFRASTRUCTURE -- not part of the product.
//
// The UNMODIFIED reference `fix eph/atomic` (/root/reference/fix_eph_atomic.cpp with eph_kappa.h, reached by include
// path, never copied) behind the same C driver as the reference `fix eph` (oracle/ref/ref_driver.cpp), built into
// oracle/_ref/libeph_atomic_ref.so.  Used to pin the restatement in oracle/eph_oracle.c (orc_atomic_*) bit for bit
// and to generate tests/golden/atomic_*.npz.
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
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
#include "eph_kappa.h"
#include "fix_eph_atomic.h"
#undef private
#undef protected

#include "fix_driver.h"

namespace {

struct RefAtomicFix : LAMMPS_NS::FixEPHAtomic {
  using LAMMPS_NS::FixEPHAtomic::FixEPHAtomic;
  // which: 0 rho[nt] 1 w[nl][3] 2 xi[nl][3] 3 f_EPH[nl][3] 4 f_RNG[nl][3] 5 rho_a[nt] 6 E_a[nt] 7 dE_a[nl] 8 T_a[nl]
  void probe_copy(int which, size_t nl, size_t nt, double *out) {
    switch (which) {
      case 0: std::copy(rho_i, rho_i + nt, out); break;
      case 1: std::copy(&w_i[0][0], &w_i[0][0] + 3 * nl, out); break;
      case 2: std::copy(&xi_i[0][0], &xi_i[0][0] + 3 * nl, out); break;
      case 3: std::copy(&f_EPH[0][0], &f_EPH[0][0] + 3 * nl, out); break;
      case 4: std::copy(&f_RNG[0][0], &f_RNG[0][0] + 3 * nl, out); break;
      case 5: std::copy(rho_a_i, rho_a_i + nt, out); break;
      case 6: for (size_t i = 0; i < nt; ++i) out[i] = E_a_i[i][0]; break;
      case 7: std::copy(dE_a_i, dE_a_i + nl, out); break;
      case 8: std::copy(T_a_i, T_a_i + nl, out); break;
      default: throw std::runtime_error("bad probe id");
    }
  }
  size_t grid_size() { return 0; }
  void grid_T(double *) {}
};

}  // namespace

SHIM_DRIVER_DEFINE(refa, RefAtomicFix)

extern "C" {

// Overwrite the per-atom electronic energies (all local atoms), e.g. to start a heat-diffusion test from a gradient.
int refa_set_energy(void *w_, const double *E) {
  auto *w = static_cast<refa_world *>(w_);
  for (int i = 0; i < w->lmp.atom->nlocal; ++i) w->fix->E_a_i[i][0] = E[i];
  return 0;
}

void *refa_kappa_load(const char *file) {
  std::ifstream probe(file);
  if (!probe.is_open()) return nullptr;
  return new Kappa(file);
}
void refa_kappa_free(void *k) { delete static_cast<Kappa *>(k); }
// dims: n_elements, n_pairs, n_r, n_T ; scal: r_cutoff, r_cutoff_sq, T_max, inv_dr_sq, dT
void refa_kappa_info(void *k_, long long *dims, double *scal) {
  Kappa *k = static_cast<Kappa *>(k_);
  dims[0] = (long long)k->n_elements; dims[1] = (long long)k->n_pairs;
  dims[2] = (long long)k->rho_r_sq[0].c.size(); dims[3] = (long long)k->E_T_atomic[0].y.size();
  scal[0] = k->r_cutoff; scal[1] = k->r_cutoff_sq; scal[2] = k->T_max; scal[3] = k->rho_r_sq[0].inv_dx;
  scal[4] = k->E_T_atomic[0].dx;
}
void refa_kappa_name(void *k_, int e, char *out, int len) {
  std::snprintf(out, len, "%s", static_cast<Kappa *>(k_)->element_name[e].c_str());
}
// kind 0: rho(r) coefficients [n_r][4]; 1: rho(r^2) coefficients [n_r][4]; 2: E(T) running sum [n_T]; 3: K(T) of pair slot e [n_T]
void refa_kappa_table(void *k_, int kind, int e, double *out) {
  Kappa *k = static_cast<Kappa *>(k_);
  if (kind < 2) {
    const Spline &s = kind == 0 ? k->rho_r[e] : k->rho_r_sq[e];
    for (size_t i = 0; i < s.c.size(); ++i) {
      out[4 * i + 0] = s.c[i].a; out[4 * i + 1] = s.c[i].b; out[4 * i + 2] = s.c[i].c; out[4 * i + 3] = s.c[i].d;
    }
  } else {
    const auto &y = kind == 2 ? k->E_T_atomic[e].y : k->K_T_atomic[e].y;
    std::copy(y.begin(), y.end(), out);
  }
}

}  // extern "C"
