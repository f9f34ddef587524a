/* TEST INFRASTRUCTURE -- parity oracle for `fix eph/atomic`, not the product.
 *
 * Plain-C restatement of the reference's per-atom electronic-energy variant: EPH_kappa (eph_kappa.h) and the
 * per-timestep path of FixEPHAtomic (fix_eph_atomic.cpp).  Every function cites the lines it follows.  Pinned bit for
 * bit against the UNMODIFIED reference compiled into oracle/_ref/libeph_atomic_ref.so
 * (tests/test_oracle_vs_reference.py::test_atomic_*).  Same arithmetic order as the reference, no FMA contraction.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "eph_oracle.h"

#define NEIGHMASK 0x1FFFFFFF

/* =========================================================================
 * EPH_kappa -- eph_kappa.h:53-151
 * ========================================================================= */
void orc_kappa_free(orc_kappa *k) {
  if (!k) return;
  free(k->rho_r); free(k->rho_r_sq); free(k->E_T); free(k->K_T); free(k);
}

orc_kappa *orc_kappa_load(const char *file) {
  FILE *fd = fopen(file, "r");
  char line[1024];
  int n_el, e;
  unsigned long n_r, n_T;
  double dr, r_cutoff, dT, T_max, dr_sq, *l_rho, *l_C;
  orc_kappa *k;
  char *tok;
  size_t j, p;
  if (!fd) return NULL;
  /* three comment lines, eph_kappa.h:60-62 */
  if (!fgets(line, sizeof line, fd) || !fgets(line, sizeof line, fd) || !fgets(line, sizeof line, fd)) goto bad;
  if (fscanf(fd, "%d", &n_el) != 1 || n_el < 1 || n_el > 16) goto bad; /* :65 */
  if (!fgets(line, sizeof line, fd)) goto bad;                         /* element names, :80-88 */
  if (fscanf(fd, "%lu %lf %lf %lu %lf %lf", &n_r, &dr, &r_cutoff, &n_T, &dT, &T_max) != 6) goto bad; /* :96-102 */
  k = (orc_kappa *)calloc(1, sizeof(orc_kappa));
  k->n_elements = n_el;
  k->n_pairs = (n_el > 1) ? (n_el + 1) * (n_el - 1) / 2 : 1;           /* :69 (sic) */
  k->n_r = n_r; k->n_T = n_T; k->dT = dT; k->T_max = T_max;
  k->r_cutoff = r_cutoff;
  k->r_cutoff_sq = r_cutoff * r_cutoff;                                /* :104 */
  dr_sq = k->r_cutoff_sq / ((double)(n_r - 1));                        /* :105 */
  k->inv_dr = 1. / dr; k->inv_dr_sq = 1. / dr_sq;                      /* Spline ctor, eph_spline.h:33 */
  k->rho_r = (double *)malloc(sizeof(double) * 4 * n_r * n_el);
  k->rho_r_sq = (double *)malloc(sizeof(double) * 4 * n_r * n_el);
  k->E_T = (double *)malloc(sizeof(double) * n_T * n_el);
  k->K_T = (double *)malloc(sizeof(double) * n_T * k->n_pairs);
  tok = strtok(line, " \t\r\n");
  for (e = 0; e < n_el; ++e) {
    snprintf(k->names[e], sizeof k->names[e], "%s", tok ? tok : "");
    tok = tok ? strtok(NULL, " \t\r\n") : NULL;
  }
  l_rho = (double *)malloc(sizeof(double) * n_r);
  l_C = (double *)malloc(sizeof(double) * n_T);
  for (e = 0; e < n_el; ++e) { /* :112-142 */
    int z, c;
    double *rho_r = k->rho_r + 4 * n_r * e;
    c = fscanf(fd, "%d", &z);
    k->number[e] = z;
    for (j = 0; j != n_r; ++j) c += fscanf(fd, "%lf", &l_rho[j]);
    orc_spline_build(dr, l_rho, n_r, rho_r);                                               /* :122 */
    for (j = 0; j < n_r; ++j) l_rho[j] = orc_spline_eval(rho_r, k->inv_dr, sqrt(j * dr_sq)); /* :125-127 */
    orc_spline_build(dr_sq, l_rho, n_r, k->rho_r_sq + 4 * n_r * e);                        /* :128 */
    for (j = 0; j != n_T; ++j) c += fscanf(fd, "%lf", &l_C[j]);                            /* :131-133 */
    if (c != (int)(1 + n_r + n_T)) { free(l_rho); free(l_C); orc_kappa_free(k); goto bad; }
    l_C[0] = 0.;                                                                           /* E(T) running sum, :136-140 */
    for (j = 1; j < n_T; ++j) l_C[j] = l_C[j - 1] + l_C[j] * dT;
    memcpy(k->E_T + n_T * e, l_C, sizeof(double) * n_T);
  }
  for (p = 0; p < (size_t)k->n_pairs; ++p)                                                 /* :145-151 */
    for (j = 0; j < n_T; ++j)
      if (fscanf(fd, "%lf", &k->K_T[n_T * p + j]) != 1) { free(l_rho); free(l_C); orc_kappa_free(k); goto bad; }
  free(l_rho); free(l_C);
  fclose(fd);
  return k;
bad:
  fclose(fd);
  return NULL;
}

void orc_kappa_info(const orc_kappa *k, long long *dims, double *scal) {
  dims[0] = k->n_elements; dims[1] = k->n_pairs; dims[2] = (long long)k->n_r; dims[3] = (long long)k->n_T;
  scal[0] = k->r_cutoff; scal[1] = k->r_cutoff_sq; scal[2] = k->T_max; scal[3] = k->inv_dr_sq; scal[4] = k->dT;
}
const double *orc_kappa_table(const orc_kappa *k, int kind, int e) {
  switch (kind) {
    case 0: return k->rho_r + 4 * k->n_r * e;
    case 1: return k->rho_r_sq + 4 * k->n_r * e;
    case 2: return k->E_T + k->n_T * e;
    default: return k->K_T + k->n_T * e;
  }
}

/* the four look-ups the fix makes */
static double kap_rho_r_sq(const orc_kappa *k, int e, double r_sq) { /* kappa.rho_r_sq[e](r_sq) */
  return orc_spline_eval(k->rho_r_sq + 4 * k->n_r * e, k->inv_dr_sq, r_sq);
}
static double kap_E_of_T(const orc_kappa *k, int e, double T) {      /* kappa.E_T_atomic[e](T), eph_linear.h:40-47 */
  return orc_linear_eval(k->dT, k->E_T + k->n_T * e, k->n_T, T);
}
static double kap_T_of_E(const orc_kappa *k, int e, double E) {      /* kappa.E_T_atomic[e].reverse(E), eph_linear.h:50-64 */
  return orc_linear_reverse(k->dT, k->E_T + k->n_T * e, k->n_T, E);
}
static double kap_K_of_T(const orc_kappa *k, int e, double T) {      /* kappa.K_T_atomic[e](T): indexed by ELEMENT, fix_eph_atomic.cpp:731 */
  return orc_linear_eval(k->dT, k->K_T + k->n_T * e, k->n_T, T);
}

/* =========================================================================
 * FixEPHAtomic -- fix_eph_atomic.cpp
 * ========================================================================= */
static void afix_resize(orc_afix *fx, size_t n) { /* grow_arrays, fix_eph_atomic.cpp:816-837; zeroed as in :108-123 */
  if (n <= fx->cap) return;
#define GROW(p, w) do { p = (double *)realloc(p, sizeof(double) * (w) * n); \
    memset(p + (w) * fx->cap, 0, sizeof(double) * (w) * (n - fx->cap)); } while (0)
  GROW(fx->rho_i, 1); GROW(fx->w_i, 3); GROW(fx->xi_i, 3); GROW(fx->f_EPH, 3); GROW(fx->f_RNG, 3); GROW(fx->array, 12);
  GROW(fx->rho_a_i, 1); GROW(fx->E_a_i, 2); GROW(fx->dE_a_i, 1); GROW(fx->T_a_i, 1);
#undef GROW
  fx->cap = n;
}

static double diff_sq(const double *x, const double *y, double *z) { /* fix_eph_atomic.h:176-182 */
  z[0] = x[0] - y[0]; z[1] = x[1] - y[1]; z[2] = x[2] - y[2];
  return z[0] * z[0] + z[1] * z[1] + z[2] * z[2];
}
static double dot3(const double *x, const double *y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }

/* Comm::forward_comm(Fix*) on one rank with periodic images (pack/unpack, fix_eph_atomic.cpp:849-927);
 * stride = row width of the array, col/width = the columns sent */
static void ghost_fill(const orc_atoms *a, double *arr, int stride, int width) {
  int g, d;
  for (g = 0; g < a->nghost; ++g)
    for (d = 0; d < width; ++d) arr[(size_t)(a->nlocal + g) * stride + d] = arr[(size_t)a->ghost_owner[g] * stride + d];
}

static void populate_array(orc_afix *fx, const orc_atoms *a) { /* fix_eph_atomic.cpp:401-435 */
  int i, c;
  for (i = 0; i < a->nlocal; ++i) {
    double *r = fx->array + 12 * (size_t)i;
    if (a->mask[i] & fx->groupbit) {
      r[0] = fx->rho_i[i];
      r[1] = orc_beta_beta(fx->beta, fx->type_map_beta[a->type[i] - 1], fx->rho_i[i]);
      for (c = 0; c < 3; ++c) { r[2 + c] = fx->f_EPH[3 * (size_t)i + c]; r[5 + c] = fx->f_RNG[3 * (size_t)i + c]; }
      r[8] = fx->rho_a_i[i];
      r[9] = fx->E_a_i[2 * (size_t)i];
      r[10] = fx->dE_a_i[i];
      r[11] = fx->T_a_i[i];
    } else {
      for (c = 0; c < 12; ++c) r[c] = 0.0;
    }
  }
}

/* the per-atom part of the constructor, fix_eph_atomic.cpp:212-257 */
orc_afix *orc_afix_new(int flags, int groupbit, int ntypes, const int *type_map_beta, const int *type_map_kappa, double dt,
                       double boltz, double ftm2v, int inner_loops, double T_init, const orc_beta *beta,
                       const orc_kappa *kappa, const orc_atoms *a) {
  orc_afix *fx = (orc_afix *)calloc(1, sizeof(orc_afix));
  int i, counter = 0;
  fx->flags = flags; fx->groupbit = groupbit; fx->ntypes = ntypes;
  for (i = 0; i < ntypes && i < 16; ++i) { fx->type_map_beta[i] = type_map_beta[i]; fx->type_map_kappa[i] = type_map_kappa[i]; }
  fx->boltz = boltz; fx->ftm2v = ftm2v;
  fx->beta = beta; fx->kappa = kappa;
  fx->inner_loops = inner_loops < 1 ? 0 : inner_loops; /* :160-164 */
  orc_afix_set_dt(fx, dt);
  afix_resize(fx, (size_t)a->nlocal + a->nghost);
  for (i = 0; i < a->nlocal; ++i)                      /* :215-221 */
    if (a->mask[i] & groupbit) fx->E_a_i[2 * (size_t)i] = kap_E_of_T(kappa, fx->type_map_kappa[a->type[i] - 1], T_init);
  fx->Ee = 0.;                                         /* :227-234 */
  for (i = 0; i < a->nlocal; ++i)
    if (a->mask[i] & groupbit) fx->Ee += fx->E_a_i[2 * (size_t)i];
  fx->Te = 0.0;                                        /* :236-253 */
  for (i = 0; i < a->nlocal; ++i)
    if (a->mask[i] & groupbit) {
      fx->T_a_i[i] = kap_T_of_E(kappa, fx->type_map_kappa[a->type[i] - 1], fx->E_a_i[2 * (size_t)i]);
      fx->Te += fx->T_a_i[i];
      counter++;
    }
  if (counter > 0) fx->Te /= (double)counter;
  fx->Te /= (double)(counter > 0 ? 1 : 0);
  populate_array(fx, a);
  return fx;
}

void orc_afix_free(orc_afix *fx) {
  if (!fx) return;
  free(fx->rho_i); free(fx->w_i); free(fx->xi_i); free(fx->f_EPH); free(fx->f_RNG); free(fx->array);
  free(fx->rho_a_i); free(fx->E_a_i); free(fx->dE_a_i); free(fx->T_a_i); free(fx);
}

void orc_afix_set_dt(orc_afix *fx, double dt) { /* fix_eph_atomic.cpp:809-814 */
  fx->dt = dt;
  fx->eta_factor = sqrt(2.0 * fx->boltz / dt);
}

void orc_atomic_calculate_environment(orc_afix *fx, const orc_atoms *a) { /* fix_eph_atomic.cpp:437-490 */
  const double rc2 = fx->beta->r_cutoff_sq, rk2 = fx->kappa->r_cutoff_sq;
  int i;
  for (i = 0; i != a->nlocal; ++i) {
    fx->rho_i[i] = 0;
    fx->rho_a_i[i] = 0;
    if (a->mask[i] & fx->groupbit) {
      long long j;
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype;
        double d[3], r_sq;
        if (!(a->mask[jj] & fx->groupbit)) continue; /* :467 */
        jtype = a->type[jj];
        r_sq = diff_sq(a->x + 3 * (size_t)jj, a->x + 3 * (size_t)i, d);
        if (r_sq < rc2) fx->rho_i[i] += orc_beta_rho_r_sq(fx->beta, fx->type_map_beta[jtype - 1], r_sq);
        if (r_sq < rk2) fx->rho_a_i[i] += kap_rho_r_sq(fx->kappa, fx->type_map_kappa[jtype - 1], r_sq);
      }
    }
  }
}

void orc_atomic_force_prl(orc_afix *fx, const orc_atoms *a) { /* fix_eph_atomic.cpp:492-679 */
  const orc_beta *b = fx->beta;
  const orc_kappa *kp = fx->kappa;
  const double rc2 = b->r_cutoff_sq;
  const double *x = a->x, *v = a->v;
  double *rho_i = fx->rho_i, *w_i = fx->w_i, *xi_i = fx->xi_i, *dE = fx->dE_a_i;
  const double l_dt = fx->dt;
  int i;
  if (fx->flags & ORC_FRICTION) {
    for (i = 0; i != a->nlocal; ++i) { /* :508-547 */
      long long j;
      double alpha_i;
      if (!(a->mask[i] & fx->groupbit)) continue;
      if (!(rho_i[i] > 0)) continue;
      alpha_i = orc_beta_alpha(b, fx->type_map_beta[a->type[i] - 1], rho_i[i]);
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype = a->type[jj];
        double e_ij[3], e_r_sq, v_rho_ji, prescaler, var1, var2, dvar;
        if (!(a->mask[jj] & fx->groupbit)) continue;
        e_r_sq = diff_sq(x + 3 * (size_t)jj, x + 3 * (size_t)i, e_ij);
        if (e_r_sq >= rc2) continue;
        v_rho_ji = orc_beta_rho_r_sq(b, fx->type_map_beta[jtype - 1], e_r_sq);
        prescaler = alpha_i * v_rho_ji / (rho_i[i] * e_r_sq);
        var1 = prescaler * dot3(e_ij, v + 3 * (size_t)i);
        var2 = prescaler * dot3(e_ij, v + 3 * (size_t)jj);
        dvar = var1 - var2;
        w_i[3 * (size_t)i + 0] += dvar * e_ij[0];
        w_i[3 * (size_t)i + 1] += dvar * e_ij[1];
        w_i[3 * (size_t)i + 2] += dvar * e_ij[2];
      }
    }
    ghost_fill(a, w_i, 3, 3); /* FixState::WI, :549-550 */
    for (i = 0; i != a->nlocal; ++i) { /* :554-606 */
      long long j;
      int itype = a->type[i];
      double alpha_i;
      if (!(a->mask[i] & fx->groupbit)) continue;
      if (!(rho_i[i] > 0)) continue;
      alpha_i = orc_beta_alpha(b, fx->type_map_beta[itype - 1], rho_i[i]);
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype = a->type[jj];
        double e_ij[3], e_r_sq, alpha_j, v_rho_ji, v_rho_ij, var1, var2, dvar, f_ij[3];
        if (!(a->mask[jj] & fx->groupbit)) continue;
        e_r_sq = diff_sq(x + 3 * (size_t)jj, x + 3 * (size_t)i, e_ij);
        if (e_r_sq >= rc2 || !(rho_i[jj] > 0)) continue;
        alpha_j = orc_beta_alpha(b, fx->type_map_beta[jtype - 1], rho_i[jj]);
        v_rho_ji = orc_beta_rho_r_sq(b, fx->type_map_beta[jtype - 1], e_r_sq);
        var1 = alpha_i * v_rho_ji * dot3(e_ij, w_i + 3 * (size_t)i) / (rho_i[i] * e_r_sq);
        v_rho_ij = orc_beta_rho_r_sq(b, fx->type_map_beta[itype - 1], e_r_sq);
        var2 = alpha_j * v_rho_ij * dot3(e_ij, w_i + 3 * (size_t)jj) / (rho_i[jj] * e_r_sq);
        dvar = var1 - var2;
        f_ij[0] = dvar * e_ij[0]; f_ij[1] = dvar * e_ij[1]; f_ij[2] = dvar * e_ij[2];
        fx->f_EPH[3 * (size_t)i + 0] -= f_ij[0];
        fx->f_EPH[3 * (size_t)i + 1] -= f_ij[1];
        fx->f_EPH[3 * (size_t)i + 2] -= f_ij[2];
        if (!(fx->flags & ORC_NOFRICTION)) { /* :599-603 */
          dE[i] += 0.5 * f_ij[0] * (v[3 * (size_t)i + 0] - v[3 * (size_t)jj + 0]) * l_dt;
          dE[i] += 0.5 * f_ij[1] * (v[3 * (size_t)i + 1] - v[3 * (size_t)jj + 1]) * l_dt;
          dE[i] += 0.5 * f_ij[2] * (v[3 * (size_t)i + 2] - v[3 * (size_t)jj + 2]) * l_dt;
        }
      }
    }
  }
  if (fx->flags & ORC_RANDOM) { /* :610-678 */
    for (i = 0; i != a->nlocal; i++) {
      long long j;
      int itype = a->type[i];
      double alpha_i, v_Ti;
      if (!(a->mask[i] & fx->groupbit)) continue;
      if (!(rho_i[i] > 0)) continue;
      alpha_i = orc_beta_alpha(b, fx->type_map_beta[itype - 1], rho_i[i]);
      v_Ti = sqrt(kap_T_of_E(kp, fx->type_map_kappa[itype - 1], fx->E_a_i[2 * (size_t)i]));
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype = a->type[jj];
        double e_ij[3], e_r_sq, alpha_j, v_Tj, v_rho_ji, v_rho_ij, var1, var2, dvar, f_ij[3];
        if (!(a->mask[jj] & fx->groupbit)) continue;
        e_r_sq = diff_sq(x + 3 * (size_t)jj, x + 3 * (size_t)i, e_ij);
        if ((e_r_sq >= rc2) || !(rho_i[jj] > 0)) continue;
        alpha_j = orc_beta_alpha(b, fx->type_map_beta[jtype - 1], rho_i[jj]);
        v_Tj = sqrt(kap_T_of_E(kp, fx->type_map_kappa[jtype - 1], fx->E_a_i[2 * (size_t)jj]));
        v_rho_ji = orc_beta_rho_r_sq(b, fx->type_map_beta[jtype - 1], e_r_sq);
        var1 = v_Ti * alpha_i * v_rho_ji * dot3(e_ij, xi_i + 3 * (size_t)i) / (rho_i[i] * e_r_sq);
        v_rho_ij = orc_beta_rho_r_sq(b, fx->type_map_beta[itype - 1], e_r_sq);
        var2 = v_Tj * alpha_j * v_rho_ij * dot3(e_ij, xi_i + 3 * (size_t)jj) / (rho_i[jj] * e_r_sq);
        dvar = fx->eta_factor * (var1 - var2);
        f_ij[0] = dvar * e_ij[0]; f_ij[1] = dvar * e_ij[1]; f_ij[2] = dvar * e_ij[2];
        fx->f_RNG[3 * (size_t)i + 0] += f_ij[0];
        fx->f_RNG[3 * (size_t)i + 1] += f_ij[1];
        fx->f_RNG[3 * (size_t)i + 2] += f_ij[2];
        if (!(fx->flags & ORC_NORANDOM)) { /* :671-675 */
          dE[i] -= 0.5 * f_ij[0] * (v[3 * (size_t)i + 0] - v[3 * (size_t)jj + 0]) * l_dt;
          dE[i] -= 0.5 * f_ij[1] * (v[3 * (size_t)i + 1] - v[3 * (size_t)jj + 1]) * l_dt;
          dE[i] -= 0.5 * f_ij[2] * (v[3 * (size_t)i + 2] - v[3 * (size_t)jj + 2]) * l_dt;
        }
      }
    }
  }
}

void orc_atomic_heat_solve(orc_afix *fx, const orc_atoms *a) { /* fix_eph_atomic.cpp:681-787 */
  const orc_kappa *kp = fx->kappa;
  const double *x = a->x;
  double *E = fx->E_a_i;
  int loops = fx->inner_loops > 0 ? fx->inner_loops : 1; /* :695-699 */
  double scaling = 1.0 / (double)loops;
  double dt = fx->dt * scaling;
  int it, j;
  for (it = 0; it < loops; ++it) {
    for (j = 0; j < a->nlocal; ++j)                        /* :705-718 */
      if (a->mask[j] & fx->groupbit) {
        E[2 * (size_t)j] += fx->dE_a_i[j] * scaling;
        if (E[2 * (size_t)j] < 0.0) E[2 * (size_t)j] = 0.0;
      }
    ghost_fill(a, E, 2, 1);                                /* FixState::EI, :720-721 */
    for (j = 0; j < a->nlocal; ++j) {                      /* :725-780 */
      E[2 * (size_t)j + 1] = E[2 * (size_t)j];
      if (a->mask[j] & fx->groupbit) {
        int jtype = a->type[j];
        long long k;
        double l_dE_j = 0.;
        double l_T_j = kap_T_of_E(kp, fx->type_map_kappa[jtype - 1], E[2 * (size_t)j]);
        double l_K_j = kap_K_of_T(kp, fx->type_map_kappa[jtype - 1], l_T_j);
        const double rho_j = fx->rho_a_i[j];
        const double rho_j_inv = 1. / fx->rho_a_i[j];
        for (k = a->offsets[j]; k != a->offsets[j + 1]; ++k) {
          int kk = a->neigh[k] & NEIGHMASK;
          int ktype = a->type[kk];
          double rho_k, rho_k_inv, l_T_k, l_K_k, l_K, v_dT, e_jk[3], e_r_sq, v_rho_j, v_rho_k;
          if (!(a->mask[kk] & fx->groupbit)) continue;
          rho_k = fx->rho_a_i[kk];
          rho_k_inv = 1. / fx->rho_a_i[kk];
          l_T_k = kap_T_of_E(kp, fx->type_map_kappa[ktype - 1], E[2 * (size_t)kk]);
          l_K_k = kap_K_of_T(kp, fx->type_map_kappa[ktype - 1], l_T_k);
          l_K = 0.5 * (l_K_j + l_K_k);
          v_dT = l_T_k - l_T_j;
          e_r_sq = diff_sq(x + 3 * (size_t)kk, x + 3 * (size_t)j, e_jk);
          if (e_r_sq >= kp->r_cutoff_sq) continue;
          v_rho_j = kap_rho_r_sq(kp, fx->type_map_kappa[jtype - 1], e_r_sq);
          v_rho_k = kap_rho_r_sq(kp, fx->type_map_kappa[ktype - 1], e_r_sq);
          if (rho_j > 0.) l_dE_j += l_K * v_rho_k * rho_j_inv * v_dT;
          if (rho_k > 0.) l_dE_j += l_K * v_rho_j * rho_k_inv * v_dT;
        }
        E[2 * (size_t)j + 1] = E[2 * (size_t)j] + 0.5 * l_dE_j * dt;
        if (E[2 * (size_t)j + 1] < 0) E[2 * (size_t)j + 1] = 0.0;
      }
    }
    for (j = 0; j < a->nlocal; ++j) E[2 * (size_t)j] = E[2 * (size_t)j + 1]; /* :783-785 */
  }
}

void orc_atomic_post_force(orc_afix *fx, const orc_atoms *a, const double *xi) { /* fix_eph_atomic.cpp:789-857 */
  int i, d;
  size_t nl = (size_t)a->nlocal;
  afix_resize(fx, nl + a->nghost);
  memset(fx->w_i, 0, sizeof(double) * 3 * nl);   /* :796-800 */
  memset(fx->xi_i, 0, sizeof(double) * 3 * nl);
  memset(fx->f_EPH, 0, sizeof(double) * 3 * nl);
  memset(fx->f_RNG, 0, sizeof(double) * 3 * nl);
  memset(fx->dE_a_i, 0, sizeof(double) * nl);
  ghost_fill(a, fx->E_a_i, 2, 1);                /* FixState::EI, :803-804 */
  if (fx->flags & ORC_RANDOM) {                  /* :808-819 */
    for (i = 0; i < a->nlocal; ++i)
      if (a->mask[i] & fx->groupbit)
        for (d = 0; d < 3; ++d) fx->xi_i[3 * (size_t)i + d] = xi[3 * (size_t)i + d];
    ghost_fill(a, fx->xi_i, 3, 3);
  }
  orc_atomic_calculate_environment(fx, a);       /* :822 */
  ghost_fill(a, fx->rho_i, 1, 1);                /* FixState::RHO carries rho_i and rho_a_i, :824-825, :855-859 */
  ghost_fill(a, fx->rho_a_i, 1, 1);
  orc_atomic_force_prl(fx, a);                   /* :827 */
  if ((fx->flags & ORC_FRICTION) && !(fx->flags & ORC_NOFRICTION)) /* :830-838: group atoms only */
    for (i = 0; i < a->nlocal; i++)
      if (a->mask[i] & fx->groupbit)
        for (d = 0; d < 3; ++d) a->f[3 * (size_t)i + d] += fx->f_EPH[3 * (size_t)i + d];
  if ((fx->flags & ORC_RANDOM) && !(fx->flags & ORC_NORANDOM))     /* :840-848 */
    for (i = 0; i < a->nlocal; i++)
      if (a->mask[i] & fx->groupbit)
        for (d = 0; d < 3; ++d) a->f[3 * (size_t)i + d] += fx->f_RNG[3 * (size_t)i + d];
}

void orc_atomic_end_of_step(orc_afix *fx, const orc_atoms *a) { /* fix_eph_atomic.cpp:361-399 */
  double E_local = 0.0, T_local = 0.0;
  int i, counter = 0;
  if (fx->flags & ORC_FDM) orc_atomic_heat_solve(fx, a); /* Flag::HEAT = 4 */
  for (i = 0; i < a->nlocal; ++i)
    if (a->mask[i] & fx->groupbit) {
      E_local += fx->E_a_i[2 * (size_t)i];
      fx->T_a_i[i] = kap_T_of_E(fx->kappa, fx->type_map_kappa[a->type[i] - 1], fx->E_a_i[2 * (size_t)i]);
      T_local += fx->T_a_i[i];
      counter++;
    }
  if (counter > 0) T_local /= (double)counter;
  fx->Ee = E_local;
  fx->Te = T_local / (double)(counter > 0 ? 1 : 0);
  populate_array(fx, a);
}

void orc_atomic_initial_integrate(orc_afix *fx, const orc_atoms *a, const double *mass) { /* fix_eph_atomic.cpp:313-337 */
  double dtv = fx->dt, dtf = 0.5 * fx->dt * fx->ftm2v;
  int i, d;
  if (fx->flags & ORC_NOINT) return;
  for (i = 0; i < a->nlocal; ++i)
    if (a->mask[i] & fx->groupbit) {
      double dtfm = dtf / mass[a->type[i]];
      for (d = 0; d < 3; ++d) a->v[3 * (size_t)i + d] += dtfm * a->f[3 * (size_t)i + d];
      for (d = 0; d < 3; ++d) a->x[3 * (size_t)i + d] += dtv * a->v[3 * (size_t)i + d];
    }
}

void orc_atomic_final_integrate(orc_afix *fx, const orc_atoms *a, const double *mass) { /* fix_eph_atomic.cpp:339-359 */
  double dtf = 0.5 * fx->dt * fx->ftm2v;
  int i, d;
  if (fx->flags & ORC_NOINT) return;
  for (i = 0; i < a->nlocal; ++i)
    if (a->mask[i] & fx->groupbit) {
      double dtfm = dtf / mass[a->type[i]];
      for (d = 0; d < 3; ++d) a->v[3 * (size_t)i + d] += dtfm * a->f[3 * (size_t)i + d];
    }
}

double *orc_afix_ptr(orc_afix *fx, int which) {
  switch (which) {
    case 0: return fx->rho_i; case 1: return fx->w_i; case 2: return fx->xi_i; case 3: return fx->f_EPH;
    case 4: return fx->f_RNG; case 5: return fx->rho_a_i; case 6: return fx->E_a_i; case 7: return fx->dE_a_i;
    case 8: return fx->T_a_i; default: return fx->array;
  }
}
double orc_afix_Ee(const orc_afix *fx) { return fx->Ee; }
double orc_afix_Te(const orc_afix *fx) { return fx->Te; }
