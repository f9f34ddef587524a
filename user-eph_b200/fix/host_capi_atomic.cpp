// C accessors over the host side of `fix eph/atomic/b200` for the Python tests: the `.kappa` tables built by
// eph_kappa_tables.h and the shim-driven FixEPHAtomicB200 (driver prefix `b200a`).  Built into
// libeph_b200_atomic_fix.so, which links libeph_b200.so.
#include <cstring>
#include <string>

#include "eph_kappa_tables.h"
#include "fix_eph_atomic_b200.h"

#include "fix_driver.h"

SHIM_DRIVER_DEFINE(b200a, LAMMPS_NS::FixEPHAtomicB200)

namespace {
thread_local std::string g_err;
}

extern "C" {

int b200a_set_energy(void *w_, const double *E) {
  auto *w = static_cast<b200a_world *>(w_);
  return shim_driver::guarded(w, [&] { w->fix->set_energy_host(E); });
}

const char *ephk_last_error(void) { return g_err.c_str(); }

void *ephk_kappa_load(const char *path) {
  try {
    return new eph_b200::KappaTables(eph_b200::load_kappa_file(path));
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void ephk_kappa_free(void *k) { delete static_cast<eph_b200::KappaTables *>(k); }
// dims: n_elements, n_pairs, n_r, n_T ; scal: r_cutoff, r_cutoff_sq, T_max, inv_dr_sq, dT
void ephk_kappa_info(void *k_, long long *dims, double *scal) {
  auto *k = static_cast<eph_b200::KappaTables *>(k_);
  dims[0] = k->n_elements; dims[1] = k->n_pairs; dims[2] = (long long)k->n_r; dims[3] = (long long)k->n_T;
  scal[0] = k->r_cutoff; scal[1] = k->r_cutoff_sq; scal[2] = k->T_max; scal[3] = k->rho_r_sq.at(0).inv_dx; scal[4] = k->dT;
}
void ephk_kappa_name(void *k_, int e, char *out, int len) {
  std::snprintf(out, len, "%s", static_cast<eph_b200::KappaTables *>(k_)->element_name.at(e).c_str());
}
// kind 0: rho(r) [n_r][4]; 1: rho(r^2) [n_r][4]; 2: E(T) of element e [n_T]; 3: K(T) of slot e [n_T]
void ephk_kappa_table(void *k_, int kind, int e, double *out) {
  auto *k = static_cast<eph_b200::KappaTables *>(k_);
  if (kind < 2) {
    const auto &t = kind == 0 ? k->rho_r.at(e) : k->rho_r_sq.at(e);
    std::memcpy(out, t.k.data(), t.k.size() * sizeof(double));
  } else {
    const auto &y = kind == 2 ? k->E_T.at(e).y : k->K_T.at(e).y;
    std::memcpy(out, y.data(), y.size() * sizeof(double));
  }
}
// the two EPH_Linear look-ups on the host (constructor path)
double ephk_linear(void *k_, int e, double x, int reverse) {
  auto *k = static_cast<eph_b200::KappaTables *>(k_);
  return reverse ? eph_b200::linear_reverse(k->E_T.at(e), x) : eph_b200::linear_eval(k->E_T.at(e), x);
}

}  // extern "C"
