// TEST INFRASTRUCTURE -- not part of the product.
//
// Builds the UNMODIFIED reference (`/root/reference/fix_eph.cpp` and its
// header-only classes, reached by include path, never copied) into
// oracle/_ref/libeph_ref.so behind a plain C interface, so that
//   * the C restatement in oracle/eph_oracle.c can be pinned against it,
//   * golden vectors under tests/golden/ can be generated from it,
//   * bench.py can time it as the CPU baseline ("kind": "reference").
// LAMMPS and MPI are played by tests/lammps_shim.
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

// EPH_FDM keeps its grids private and the Spline its coefficients protected;
// a checker needs to read them at full precision (the writers print %.6e).
#define private public
#define protected public
#include "eph_spline.h"
#include "eph_linear.h"
#include "eph_beta.h"
#include "eph_fdm.h"
#include "fix_eph.h"
#undef private
#undef protected

#include "fix_driver.h"

namespace {

struct RefFix : LAMMPS_NS::FixEPH {
  using LAMMPS_NS::FixEPH::FixEPH;
  void probe_copy(int which, size_t nl, size_t nt, double *out) {
    switch (which) {
      case 0: std::copy(rho_i, rho_i + nt, out); break;
      case 1: std::copy(&w_i[0][0], &w_i[0][0] + 3 * nl, out); break;
      case 2: std::copy(&xi_i[0][0], &xi_i[0][0] + 3 * nl, out); break;
      case 3: std::copy(&f_EPH[0][0], &f_EPH[0][0] + 3 * nl, out); break;
      case 4: std::copy(&f_RNG[0][0], &f_RNG[0][0] + 3 * nl, out); break;
      default: throw std::runtime_error("bad probe id");
    }
  }
  size_t grid_size() { return fdm.ntotal; }
  void grid_T(double *out) { std::copy(fdm.T_e.begin(), fdm.T_e.end(), out); }
};

}  // namespace

SHIM_DRIVER_DEFINE(ref, RefFix)

// ---------------------------------------------------------------------------
// Stand-alone probes of the header-only classes (no fix, no LAMMPS).
// ---------------------------------------------------------------------------
extern "C" {

// Spline(dx, y): writes n*4 coefficients {a,b,c,d}
void ref_spline_build(double dx, const double *y, int n, double *coeff) {
  std::vector<double> yy(y, y + n);
  Spline s(dx, yy);
  for (int i = 0; i < n; ++i) {
    coeff[4 * i + 0] = s.c[i].a; coeff[4 * i + 1] = s.c[i].b;
    coeff[4 * i + 2] = s.c[i].c; coeff[4 * i + 3] = s.c[i].d;
  }
}

void ref_spline_eval(double dx, const double *y, int n, const double *x, int m, double *out) {
  std::vector<double> yy(y, y + n);
  Spline s(dx, yy);
  for (int i = 0; i < m; ++i) out[i] = s(x[i]);
}

void *ref_beta_load(const char *file) {
  std::ifstream probe(file);
  if (!probe.is_open()) return nullptr;
  return new Beta(file);
}
void ref_beta_free(void *b) { delete static_cast<Beta *>(b); }
// dims: n_elements, n_rho, n_beta ; scal: r_cutoff, r_cutoff_sq, rho_cutoff, inv_dr, inv_dr_sq, inv_drho
void ref_beta_info(void *b_, long long *dims, double *scal) {
  Beta *b = static_cast<Beta *>(b_);
  dims[0] = (long long)b->n_elements;
  dims[1] = (long long)b->rho[0].c.size();
  dims[2] = (long long)b->beta[0].c.size();
  scal[0] = b->r_cutoff; scal[1] = b->r_cutoff_sq; scal[2] = b->rho_cutoff;
  scal[3] = b->rho[0].inv_dx; scal[4] = b->rho_r_sq[0].inv_dx; scal[5] = b->beta[0].inv_dx;
}
void ref_beta_name(void *b_, int e, char *out, int len) {
  std::string s = static_cast<Beta *>(b_)->get_element_name(e);
  std::snprintf(out, len, "%s", s.c_str());
}
// kind: 0 rho(r) 1 rho(r^2) 2 alpha 3 beta
void ref_beta_table(void *b_, int kind, int e, double *coeff) {
  Beta *b = static_cast<Beta *>(b_);
  const Spline &s = kind == 0 ? b->rho[e] : kind == 1 ? b->rho_r_sq[e] : kind == 2 ? b->alpha[e] : b->beta[e];
  for (size_t i = 0; i < s.c.size(); ++i) {
    coeff[4 * i + 0] = s.c[i].a; coeff[4 * i + 1] = s.c[i].b;
    coeff[4 * i + 2] = s.c[i].c; coeff[4 * i + 3] = s.c[i].d;
  }
}
void ref_beta_eval(void *b_, int kind, int e, const double *x, int m, double *out) {
  Beta *b = static_cast<Beta *>(b_);
  for (int i = 0; i < m; ++i)
    out[i] = kind == 0 ? b->get_rho(e, x[i]) : kind == 1 ? b->get_rho_r_sq(e, x[i])
           : kind == 2 ? b->get_alpha(e, x[i]) : b->get_beta(e, x[i]);
}

// ---- EPH_FDM ----
void *ref_fdm_new(int nx, int ny, int nz, const double *box, double T_e, double C_e, double rho_e, double kappa_e) {
  auto *f = new EPH_FDM(nx, ny, nz, box[0], box[1], box[2], box[3], box[4], box[5], T_e, C_e, rho_e, kappa_e);
  f->set_comm(0, 0, 1);
  return f;
}
void *ref_fdm_from_file(const char *file) {
  std::ifstream probe(file);
  if (!probe.is_open()) return nullptr;
  auto *f = new EPH_FDM(file);
  f->set_comm(0, 0, 1);
  return f;
}
void ref_fdm_free(void *f) { delete static_cast<EPH_FDM *>(f); }
void ref_fdm_dims(void *f_, long long *d) {
  auto *f = static_cast<EPH_FDM *>(f_);
  d[0] = f->nx; d[1] = f->ny; d[2] = f->nz; d[3] = f->steps;
}
void ref_fdm_set_dt(void *f, double dt) { static_cast<EPH_FDM *>(f)->set_dt(dt); }
void ref_fdm_set_steps(void *f, long long s) { static_cast<EPH_FDM *>(f)->set_steps((size_t)s); }
// which: 0 T_e 1 S_e 2 rho_e 3 C_e 4 kappa_e 5 dT_e (double) ; flags via the *_flags calls
static std::vector<double> &fdm_field(EPH_FDM *f, int which) {
  switch (which) {
    case 0: return f->T_e; case 1: return f->S_e; case 2: return f->rho_e;
    case 3: return f->C_e; case 4: return f->kappa_e; default: return f->dT_e;
  }
}
void ref_fdm_get(void *f_, int which, double *out) {
  auto &v = fdm_field(static_cast<EPH_FDM *>(f_), which);
  std::copy(v.begin(), v.end(), out);
}
void ref_fdm_set(void *f_, int which, const double *in) {
  auto &v = fdm_field(static_cast<EPH_FDM *>(f_), which);
  std::copy(in, in + v.size(), v.begin());
}
void ref_fdm_get_flags(void *f_, short *flag, unsigned short *tdyn) {
  auto *f = static_cast<EPH_FDM *>(f_);
  std::copy(f->flag.begin(), f->flag.end(), flag);
  std::copy(f->T_dynamic_flag.begin(), f->T_dynamic_flag.end(), tdyn);
}
void ref_fdm_set_flags(void *f_, const short *flag, const unsigned short *tdyn) {
  auto *f = static_cast<EPH_FDM *>(f_);
  if (flag) std::copy(flag, flag + f->ntotal, f->flag.begin());
  if (tdyn) std::copy(tdyn, tdyn + f->ntotal, f->T_dynamic_flag.begin());
}
void ref_fdm_insert_energy(void *f_, int n, const double *x, const double *E) {
  auto *f = static_cast<EPH_FDM *>(f_);
  for (int i = 0; i < n; ++i) f->insert_energy(x[3 * i], x[3 * i + 1], x[3 * i + 2], E[i]);
}
void ref_fdm_get_T_at(void *f_, int n, const double *x, double *out) {
  auto *f = static_cast<EPH_FDM *>(f_);
  for (int i = 0; i < n; ++i) out[i] = f->get_T(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
}
void ref_fdm_index_at(void *f_, int n, const double *x, long long *out) {
  auto *f = static_cast<EPH_FDM *>(f_);
  for (int i = 0; i < n; ++i) out[i] = (long long)f->get_index(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
}
void ref_fdm_solve(void *f) { static_cast<EPH_FDM *>(f)->solve(); }
double ref_fdm_T_total(void *f) { return static_cast<EPH_FDM *>(f)->get_T_total(); }
void ref_fdm_save_temperature(void *f, const char *fn, int n) { static_cast<EPH_FDM *>(f)->save_temperature(fn, n); }
void ref_fdm_save_state(void *f, const char *fn) { static_cast<EPH_FDM *>(f)->save_state(fn); }

// ---- EPH_Linear ----
void ref_linear_eval(double dx, const double *y, int n, const double *x, int m, double *out, int reverse) {
  EPH_Linear lin(dx, y, y + n);
  for (int i = 0; i < m; ++i) out[i] = reverse ? lin.reverse_lookup(x[i]) : lin(x[i]);
}

}  // extern "C"
