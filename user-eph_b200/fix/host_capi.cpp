// C accessors over the host-side C++ of the product (table construction and
// grid-file handling) plus the shim-driven FixEPHB200, for the Python tests
// and benchmarks.  Built into libeph_b200_fix.so, which links libeph_b200.so.
#include <cstring>
#include <string>

#include "eph_grid_io.h"
#include "eph_tables.h"
#include "fix_eph_b200.h"

#include "fix_driver.h"

SHIM_DRIVER_DEFINE(b200, LAMMPS_NS::FixEPHB200)

namespace {
thread_local std::string g_err;
}

extern "C" {

const char *ephh_last_error(void) { return g_err.c_str(); }

void *ephh_beta_load(const char *path) {
  try {
    return new eph_b200::BetaTables(eph_b200::load_beta_file(path));
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
// tables from in-memory knots: rho_knots [n_el][n_rho], beta_knots [n_el][n_beta]
void *ephh_beta_from_knots(int n_el, long long n_rho, double dr, long long n_beta, double drho, double r_cutoff,
                           const double *rho_knots, const double *beta_knots) {
  try {
    auto *b = new eph_b200::BetaTables;
    b->n_elements = n_el;
    eph_b200::set_beta_header(*b, (size_t)n_rho, dr, (size_t)n_beta, drho, r_cutoff);
    for (int e = 0; e < n_el; ++e) {
      b->element_name.push_back("E" + std::to_string(e));
      b->element_number.push_back(0);
      eph_b200::build_element_tables(*b, std::vector<double>(rho_knots + e * n_rho, rho_knots + (e + 1) * n_rho),
                                     std::vector<double>(beta_knots + e * n_beta, beta_knots + (e + 1) * n_beta));
    }
    return b;
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void ephh_beta_free(void *b) { delete static_cast<eph_b200::BetaTables *>(b); }
void ephh_beta_info(void *b_, long long *dims, double *scal) {
  auto *b = static_cast<eph_b200::BetaTables *>(b_);
  dims[0] = b->n_elements; dims[1] = (long long)b->n_rho; dims[2] = (long long)b->n_beta;
  scal[0] = b->r_cutoff; scal[1] = b->r_cutoff_sq; scal[2] = b->rho_cutoff;
  scal[3] = b->rho_r[0].inv_dx; scal[4] = b->inv_dr_sq(); scal[5] = b->inv_drho();
}
void ephh_beta_name(void *b_, int e, char *out, int len) {
  std::snprintf(out, len, "%s", static_cast<eph_b200::BetaTables *>(b_)->element_name.at(e).c_str());
}
// kind: 0 rho(r) 1 rho(r^2) 2 alpha 3 beta ; coeff [n][4]
void ephh_beta_table(void *b_, int kind, int e, double *coeff) {
  auto *b = static_cast<eph_b200::BetaTables *>(b_);
  const auto &t = kind == 0 ? b->rho_r[e] : kind == 1 ? b->rho_r_sq[e] : kind == 2 ? b->alpha[e] : b->beta[e];
  std::memcpy(coeff, t.k.data(), t.k.size() * sizeof(double));
}
void ephh_spline_build(double dx, const double *y, int n, double *coeff) {
  auto t = eph_b200::make_cubic_table(dx, std::vector<double>(y, y + n));
  std::memcpy(coeff, t.k.data(), t.k.size() * sizeof(double));
}

void *ephh_grid_load(const char *path) {
  try {
    return new eph_b200::GridState(eph_b200::load_grid_file(path));
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void ephh_grid_free(void *g) { delete static_cast<eph_b200::GridState *>(g); }
void ephh_grid_dims(void *g_, long long *d, double *box) {
  auto *g = static_cast<eph_b200::GridState *>(g_);
  d[0] = g->nx; d[1] = g->ny; d[2] = g->nz; d[3] = g->steps; d[4] = g->has_tables ? (long long)g->C_e_T.size() : 0;
  std::memcpy(box, g->box, sizeof g->box);
}
// which: 0 T_e 1 S_e 2 rho_e 3 C_e 4 kappa_e
void ephh_grid_field(void *g_, int which, double *out) {
  auto *g = static_cast<eph_b200::GridState *>(g_);
  const auto &v = which == 0 ? g->T_e : which == 1 ? g->S_e : which == 2 ? g->rho_e : which == 3 ? g->C_e : g->kappa_e;
  std::memcpy(out, v.data(), v.size() * sizeof(double));
}
void ephh_grid_flags(void *g_, short *flag, unsigned short *tdyn) {
  auto *g = static_cast<eph_b200::GridState *>(g_);
  std::memcpy(flag, g->flag.data(), g->flag.size() * sizeof(short));
  std::memcpy(tdyn, g->t_dyn.data(), g->t_dyn.size() * sizeof(unsigned short));
}
// temperature tables: C_e(T), kappa_e(T) coefficients [n][4], E_e(T) [n]; returns dT
double ephh_grid_tables(void *g_, double *C, double *K, double *E) {
  auto *g = static_cast<eph_b200::GridState *>(g_);
  std::memcpy(C, g->C_e_T.k.data(), g->C_e_T.k.size() * sizeof(double));
  std::memcpy(K, g->kappa_e_T.k.data(), g->kappa_e_T.k.size() * sizeof(double));
  std::memcpy(E, g->E_e_T.y.data(), g->E_e_T.y.size() * sizeof(double));
  return g->E_e_T.dx;
}
int ephh_grid_write_heat_map(void *g_, const double *T, const char *name, int counter) {
  try {
    auto *g = static_cast<eph_b200::GridState *>(g_);
    eph_b200::write_heat_map(*g, std::vector<double>(T, T + g->ncell()), name, counter);
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int ephh_grid_write_restart(void *g_, const double *T, const char *path) {
  try {
    auto *g = static_cast<eph_b200::GridState *>(g_);
    eph_b200::write_restart(*g, std::vector<double>(T, T + g->ncell()), path);
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

}  // extern "C"
