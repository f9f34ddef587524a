/* TEST INFRASTRUCTURE -- the parity oracle, not the product (see eph_oracle.h).
 *
 * Plain-C restatement of the reference algorithm.  Arithmetic is written in
 * the reference's own operation order and this file is compiled with
 * -ffp-contract=off, so results are bit-identical to the compiled reference
 * (checked by tests/test_oracle_vs_reference.py).
 */
#include "eph_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NEIGHMASK 0x1FFFFFFF

/* =========================================================================
 * EPH_Spline -- eph_spline.h:31-131 (construction), :134-142 (evaluation)
 * ========================================================================= */
void orc_spline_build(double dx, const double *y, size_t n, double *k) {
#define A(i) k[4 * (i) + 0]
#define B(i) k[4 * (i) + 1]
#define C(i) k[4 * (i) + 2]
#define D(i) k[4 * (i) + 3]
  size_t i;
  double z0, z1, z2, z3;
  for (i = 0; i < 4 * n; ++i) k[i] = 0.0; /* vector<Coefficients>(n) value-initialises, eph_spline.h:34 */

  /* finite-difference slopes, eph_spline.h:48-51 */
  for (i = 0; i + 1 < n; ++i) B(i) = (y[i + 1] - y[i]) / dx;

  /* extrapolated slopes beyond both ends, eph_spline.h:53-59 (B(n-1) is still 0 here) */
  z1 = 2.0 * B(0) - B(1);
  z0 = 2.0 * z1 - B(0);
  z2 = 2.0 * B(n - 2) - B(n - 1);
  z3 = 2.0 * z2 - B(n - 1);
  B(n - 1) = z2;

  /* Akima weights, eph_spline.h:61-79 */
  for (i = 2; i + 2 < n; ++i) {
    C(i) = fabs(B(i + 1) - B(i));
    D(i) = fabs(B(i - 1) - B(i - 2));
  }
  C(0) = fabs(B(1) - B(0));
  D(0) = fabs(z1 - z0);
  C(1) = fabs(B(2) - B(1));
  D(1) = fabs(B(0) - z1);
  C(n - 2) = fabs(z2 - B(n - 2));
  D(n - 2) = fabs(B(n - 3) - B(n - 4));
  C(n - 1) = fabs(z3 - z2);
  D(n - 1) = fabs(B(n - 2) - B(n - 3));

  /* knot derivatives, eph_spline.h:82-108.  For the last knot the reference
   * reads c[n].b, one element past its vector (eph_spline.h:92); that value
   * only enters the equality tests below, never the arithmetic, and the tests
   * give the same branch result for any finite value, so 0 is used here. */
  for (i = 0; i < n; ++i) {
    double w0, w1, d_2, d_1, d0, d1;
    if (i == 0) { d_2 = z0; d_1 = z1; d1 = B(1); }
    else if (i == 1) { d_2 = z1; d_1 = B(0); d1 = B(2); }
    else { d_2 = B(i - 2); d_1 = B(i - 1); d1 = (i + 1 < n) ? B(i + 1) : 0.0; }
    d0 = B(i); w1 = C(i); w0 = D(i);
    if (d_2 == d_1 && d0 != d1) A(i) = d_1;
    else if (d0 == d1 && d_2 == d_1) A(i) = d0;
    else if (d_1 == d0) A(i) = d0;
    else if (d_2 == d_1 && d0 == d1 && d0 != d_1) A(i) = 0.5 * (d_1 + d0);
    else A(i) = (d_1 * w1 + d0 * w0) / (w1 + w0);
  }

  /* cubic through (y_i, y'_i), (y_i+1, y'_i+1) in ABSOLUTE x, eph_spline.h:111-125 */
  for (i = 0; i + 1 < n; ++i) {
    double dx3 = dx * dx * dx;
    double x0_1 = i * dx;
    double x0_2 = i * dx * x0_1;
    double x0_3 = i * dx * x0_2;
    double x1_1 = (i + 1) * dx;
    double x1_2 = (i + 1) * dx * x1_1;
    double x1_3 = (i + 1) * dx * x1_2;
    D(i) = (-A(i) * x0_1 - A(i + 1) * x0_1 + A(i) * x1_1 + A(i + 1) * x1_1 + 2.0 * y[i] - 2.0 * y[i + 1]) / dx3;
    C(i) = (-A(i) + A(i + 1) + 3.0 * D(i) * x0_2 - 3.0 * D(i) * x1_2) / 2.0 / dx;
    B(i) = (C(i) * x0_2 + D(i) * x0_3 - C(i) * x1_2 - D(i) * x1_3 - y[i] + y[i + 1]) / dx;
    A(i) = y[i] - B(i) * x0_1 - C(i) * x0_2 - D(i) * x0_3;
  }
  A(n - 1) = y[n - 1]; /* eph_spline.h:127-130 */
  B(n - 1) = 0.0;
  C(n - 1) = 0.0;
  D(n - 1) = 0.0;
#undef A
#undef B
#undef C
#undef D
}

/* eph_spline.h:134-142: truncating index, Horner in absolute x, no bounds check */
double orc_spline_eval(const double *k, double inv_dx, double x) {
  size_t index = (size_t)(x * inv_dx);
  const double *c = k + 4 * index;
  return c[0] + x * (c[1] + x * (c[2] + x * c[3]));
}

/* =========================================================================
 * EPH_Linear -- eph_linear.h:27-62
 * ========================================================================= */
double orc_linear_eval(double dx, const double *y, size_t n, double x) {
  size_t idx = (size_t)(x / dx); /* eph_linear.h:41 */
  if (idx < n) {
    /* the reference indexes dy[idx] with idx == n-1 possible (one past dy, eph_linear.h:44);
       the harness never evaluates in the last interval */
    double dy = (idx + 1 < n) ? (y[idx + 1] - y[idx]) / dx : 0.0;
    double delta = x - idx * dx;
    return y[idx] + dy * delta;
  }
  return 0.;
}

double orc_linear_reverse(double dx, const double *y, size_t n, double yv) {
  /* std::upper_bound: first element strictly greater than yv, eph_linear.h:51 */
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = lo + (hi - lo) / 2;
    if (!(yv < y[mid])) lo = mid + 1; else hi = mid;
  }
  if (lo != n) {
    size_t idx = lo - 1; /* eph_linear.h:53-55 */
    double dy = (y[idx + 1] - y[idx]) / dx;
    return idx * dx + 1. / dy * (yv - y[idx]);
  }
  return 0.;
}

/* =========================================================================
 * EPH_Beta -- eph_beta.h:39-128 (file), :96-125 (tables), :164-198 (lookups)
 * ========================================================================= */
static orc_beta *beta_alloc(int n_el, size_t n_rho, size_t n_beta) {
  orc_beta *b = (orc_beta *)calloc(1, sizeof(orc_beta));
  b->n_elements = n_el; b->n_rho = n_rho; b->n_beta = n_beta;
  b->rho_r = (double *)malloc(sizeof(double) * 4 * n_rho * n_el);
  b->rho_r_sq = (double *)malloc(sizeof(double) * 4 * n_rho * n_el);
  b->alpha = (double *)malloc(sizeof(double) * 4 * n_beta * n_el);
  b->beta = (double *)malloc(sizeof(double) * 4 * n_beta * n_el);
  return b;
}

static void beta_build_element(orc_beta *b, int e, double *l_rho, double *l_beta) {
  size_t j;
  double *rho_r = b->rho_r + 4 * b->n_rho * e;
  orc_spline_build(b->dr, l_rho, b->n_rho, rho_r); /* eph_beta.h:105 */
  for (j = 0; j != b->n_rho; ++j)                   /* rho(r) resampled on an r^2 grid, eph_beta.h:108-109 */
    l_rho[j] = orc_spline_eval(rho_r, b->inv_dr, sqrt(j * b->dr_sq));
  orc_spline_build(b->dr_sq, l_rho, b->n_rho, b->rho_r_sq + 4 * b->n_rho * e); /* :111 */
  orc_spline_build(b->drho, l_beta, b->n_beta, b->beta + 4 * b->n_beta * e);   /* :117 */
  for (j = 0; j != b->n_beta; ++j) l_beta[j] = sqrt(l_beta[j]);                /* :120-121 */
  orc_spline_build(b->drho, l_beta, b->n_beta, b->alpha + 4 * b->n_beta * e);  /* :123 */
}

static void beta_set_scalars(orc_beta *b, double dr, double drho, double r_cutoff) {
  b->dr = dr; b->drho = drho; b->r_cutoff = r_cutoff;
  b->r_cutoff_sq = r_cutoff * r_cutoff;          /* eph_beta.h:90 */
  b->rho_cutoff = drho * (b->n_beta - 1);        /* :91 */
  b->dr_sq = b->r_cutoff_sq / (b->n_rho - 1);    /* :93 */
  b->inv_dr = 1. / dr;                           /* Spline ctor, eph_spline.h:33 */
  b->inv_dr_sq = 1. / b->dr_sq;
  b->inv_drho = 1. / drho;
}

orc_beta *orc_beta_from_knots(int n_el, size_t n_rho, double dr, size_t n_beta, double drho, double r_cutoff,
                              const double *rho_knots, const double *beta_knots) {
  orc_beta *b = beta_alloc(n_el, n_rho, n_beta);
  double *l_rho = (double *)malloc(sizeof(double) * n_rho), *l_beta = (double *)malloc(sizeof(double) * n_beta);
  int e;
  beta_set_scalars(b, dr, drho, r_cutoff);
  for (e = 0; e < n_el; ++e) {
    memcpy(l_rho, rho_knots + n_rho * e, sizeof(double) * n_rho);
    memcpy(l_beta, beta_knots + n_beta * e, sizeof(double) * n_beta);
    snprintf(b->names[e], sizeof b->names[e], "E%d", e);
    beta_build_element(b, e, l_rho, l_beta);
  }
  free(l_rho); free(l_beta);
  return b;
}

orc_beta *orc_beta_load(const char *file) {
  FILE *fd = fopen(file, "r");
  char line[1024];
  int n_el, e, c;
  unsigned long n_rho, n_beta;
  double dr, drho, r_cutoff, *l_rho, *l_beta;
  orc_beta *b;
  char *tok;
  if (!fd) return NULL;
  /* three comment lines, eph_beta.h:46-50 */
  if (!fgets(line, sizeof line, fd) || !fgets(line, sizeof line, fd) || !fgets(line, sizeof line, fd)) goto bad;
  if (fscanf(fd, "%d", &n_el) != 1 || n_el < 1 || n_el > 16) goto bad; /* :53 */
  if (!fgets(line, sizeof line, fd)) goto bad;                         /* rest of the header line: names, :67-75 */
  if (fscanf(fd, "%lu %lf %lu %lf %lf", &n_rho, &dr, &n_beta, &drho, &r_cutoff) != 5) goto bad; /* :84-88 */
  b = beta_alloc(n_el, n_rho, n_beta);
  beta_set_scalars(b, dr, drho, r_cutoff);
  tok = strtok(line, " \t\r\n");
  for (e = 0; e < n_el; ++e) {
    snprintf(b->names[e], sizeof b->names[e], "%s", tok ? tok : "");
    tok = tok ? strtok(NULL, " \t\r\n") : NULL;
  }
  l_rho = (double *)malloc(sizeof(double) * n_rho);
  l_beta = (double *)malloc(sizeof(double) * n_beta);
  for (e = 0; e < n_el; ++e) { /* :96-125 */
    size_t j;
    unsigned z;
    c = fscanf(fd, "%u", &z);
    b->number[e] = (int)z;
    for (j = 0; j != n_rho; ++j) c += fscanf(fd, "%lf", &l_rho[j]);
    for (j = 0; j != n_beta; ++j) c += fscanf(fd, "%lf", &l_beta[j]);
    if (c != (int)(1 + n_rho + n_beta)) { free(l_rho); free(l_beta); orc_beta_free(b); goto bad; }
    beta_build_element(b, e, l_rho, l_beta);
  }
  free(l_rho); free(l_beta);
  fclose(fd);
  return b;
bad:
  fclose(fd);
  return NULL;
}

void orc_beta_free(orc_beta *b) {
  if (!b) return;
  free(b->rho_r); free(b->rho_r_sq); free(b->alpha); free(b->beta); free(b);
}

/* eph_beta.h:164-169 */
double orc_beta_rho_r_sq(const orc_beta *b, int e, double r_sq) {
  return orc_spline_eval(b->rho_r_sq + 4 * b->n_rho * e, b->inv_dr_sq, r_sq);
}
/* eph_beta.h:157-162: rho(r) of the legacy models (no cut-off test under NDEBUG) */
double orc_beta_rho_r(const orc_beta *b, int e, double r) {
  return orc_spline_eval(b->rho_r + 4 * b->n_rho * e, b->inv_dr, r);
}
/* eph_beta.h:186-198: zero above rho_cutoff */
double orc_beta_alpha(const orc_beta *b, int e, double rho) {
  if (rho > b->rho_cutoff) return 0.;
  return orc_spline_eval(b->alpha + 4 * b->n_beta * e, b->inv_drho, rho);
}
/* eph_beta.h:171-184 */
double orc_beta_beta(const orc_beta *b, int e, double rho) {
  if (rho > b->rho_cutoff) return 0.;
  return orc_spline_eval(b->beta + 4 * b->n_beta * e, b->inv_drho, rho);
}

void orc_beta_info(const orc_beta *b, long long *dims, double *scal) {
  dims[0] = b->n_elements; dims[1] = (long long)b->n_rho; dims[2] = (long long)b->n_beta;
  scal[0] = b->r_cutoff; scal[1] = b->r_cutoff_sq; scal[2] = b->rho_cutoff;
  scal[3] = b->inv_dr; scal[4] = b->inv_dr_sq; scal[5] = b->inv_drho;
}
const double *orc_beta_table(const orc_beta *b, int kind, int e) {
  switch (kind) {
    case 0: return b->rho_r + 4 * b->n_rho * e;
    case 1: return b->rho_r_sq + 4 * b->n_rho * e;
    case 2: return b->alpha + 4 * b->n_beta * e;
    default: return b->beta + 4 * b->n_beta * e;
  }
}

/* =========================================================================
 * EPH_FDM -- eph_fdm.h
 * ========================================================================= */
static void fdm_resize(orc_fdm *f, size_t nx, size_t ny, size_t nz) { /* eph_fdm.h:462-476 */
  size_t i, n = nx * ny * nz;
  f->nx = nx; f->ny = ny; f->nz = nz; f->ntotal = n;
  f->T_e = (double *)calloc(n, sizeof(double));
  f->dT_e = (double *)calloc(n, sizeof(double));
  f->ddT_e = (double *)calloc(n, sizeof(double));
  f->C_e = (double *)calloc(n, sizeof(double));
  f->rho_e = (double *)calloc(n, sizeof(double));
  f->kappa_e = (double *)calloc(n, sizeof(double));
  f->S_e = (double *)calloc(n, sizeof(double));
  f->flag = (short *)calloc(n, sizeof(short));
  f->T_dyn = (unsigned short *)calloc(n, sizeof(unsigned short));
  for (i = 0; i < n; ++i) f->flag[i] = 1;
}

static void fdm_set_box(orc_fdm *f, const double *b) { /* eph_fdm.h:122-140 */
  f->x0 = b[0]; f->x1 = b[1]; f->y0 = b[2]; f->y1 = b[3]; f->z0 = b[4]; f->z1 = b[5];
  f->dx = (b[1] - b[0]) / f->nx;
  f->dy = (b[3] - b[2]) / f->ny;
  f->dz = (b[5] - b[4]) / f->nz;
  f->dV = f->dx * f->dy * f->dz;
}

orc_fdm *orc_fdm_new(size_t nx, size_t ny, size_t nz, const double *box, double T_e, double C_e, double rho_e,
                     double kappa_e) { /* eph_fdm.h:28-46, :143-153 */
  orc_fdm *f = (orc_fdm *)calloc(1, sizeof(orc_fdm));
  size_t i;
  fdm_resize(f, nx, ny, nz);
  fdm_set_box(f, box);
  for (i = 0; i < f->ntotal; ++i) {
    f->T_e[i] = T_e; f->rho_e[i] = rho_e; f->C_e[i] = C_e; f->kappa_e[i] = kappa_e;
    f->flag[i] = 1; f->T_dyn[i] = 0;
  }
  f->steps = 1;
  f->dt = 1;
  strcpy(f->parameter_filename, "NULL");
  return f;
}

orc_fdm *orc_fdm_from_file(const char *file) { /* eph_fdm.h:48-119 */
  FILE *fd = fopen(file, "r");
  char line[1024];
  unsigned long nx, ny, nz, steps;
  double box[6];
  orc_fdm *f;
  size_t i;
  if (!fd) return NULL;
  if (!fgets(line, sizeof line, fd) || !fgets(line, sizeof line, fd) || !fgets(line, sizeof line, fd)) goto bad;
  if (fscanf(fd, "%lu %lu %lu %lu", &nx, &ny, &nz, &steps) != 4) goto bad;
  if (fscanf(fd, "%lf %lf %lf %lf %lf %lf", box, box + 1, box + 2, box + 3, box + 4, box + 5) != 6) goto bad;
  f = (orc_fdm *)calloc(1, sizeof(orc_fdm));
  fdm_resize(f, nx, ny, nz);
  f->steps = steps;
  fdm_set_box(f, box);
  if (fscanf(fd, "%1023s", f->parameter_filename) != 1) { orc_fdm_free(f); goto bad; }
  if (strcmp(f->parameter_filename, "NULL") != 0) { /* eph_fdm.h:74-104 */
    FILE *pf = fopen(f->parameter_filename, "r");
    unsigned long n;
    double dT, *C, *K;
    if (!pf) { orc_fdm_free(f); goto bad; }
    if (!fgets(line, sizeof line, pf) || !fgets(line, sizeof line, pf) || !fgets(line, sizeof line, pf) ||
        fscanf(pf, "%lu %lf", &n, &dT) != 2) { fclose(pf); orc_fdm_free(f); goto bad; }
    C = (double *)malloc(sizeof(double) * n); K = (double *)malloc(sizeof(double) * n);
    for (i = 0; i < n; ++i)
      if (fscanf(pf, "%lf %lf", &C[i], &K[i]) != 2) { fclose(pf); free(C); free(K); orc_fdm_free(f); goto bad; }
    fclose(pf);
    f->n_T = n; f->dT = dT;
    f->C_e_T = (double *)malloc(sizeof(double) * 4 * n);
    f->kappa_e_T = (double *)malloc(sizeof(double) * 4 * n);
    orc_spline_build(dT, C, n, f->C_e_T);
    orc_spline_build(dT, K, n, f->kappa_e_T);
    C[0] = 0.; /* E_e(T) = running sum of C_e dT, eph_fdm.h:98-102 */
    for (i = 1; i < n; ++i) C[i] = C[i - 1] + C[i] * dT;
    f->E_e_T = C;
    free(K);
  }
  for (i = 0; i != f->ntotal; ++i) { /* eph_fdm.h:107-118 */
    int lx, ly, lz, fl, td;
    size_t index;
    double v[5];
    if (fscanf(fd, "%d %d %d %lf %lf %lf %lf %lf %d %d", &lx, &ly, &lz, v, v + 1, v + 2, v + 3, v + 4, &fl, &td) != 10) {
      orc_fdm_free(f); goto bad;
    }
    index = lx + ly * f->nx + lz * f->nx * f->ny;
    f->T_e[index] = v[0]; f->S_e[index] = v[1]; f->rho_e[index] = v[2]; f->C_e[index] = v[3];
    f->kappa_e[index] = v[4]; f->flag[index] = (short)fl; f->T_dyn[index] = (unsigned short)td;
  }
  fclose(fd);
  return f;
bad:
  fclose(fd);
  return NULL;
}

void orc_fdm_free(orc_fdm *f) {
  if (!f) return;
  free(f->T_e); free(f->dT_e); free(f->ddT_e); free(f->C_e); free(f->rho_e); free(f->kappa_e); free(f->S_e);
  free(f->flag); free(f->T_dyn); free(f->C_e_T); free(f->kappa_e_T); free(f->E_e_T); free(f);
}

void orc_fdm_set_dt(orc_fdm *f, double dt) { f->dt = dt; }

/* eph_fdm.h:494-509: floor, then periodic wrap through a second floor */
size_t orc_fdm_index(const orc_fdm *f, double x, double y, double z) {
  int lx = (int)floor((x - f->x0) / f->dx);
  int px = (int)floor(((double)lx) / f->nx);
  int ly, py, lz, pz;
  lx -= px * (int)f->nx;
  ly = (int)floor((y - f->y0) / f->dy);
  py = (int)floor(((double)ly) / f->ny);
  ly -= py * (int)f->ny;
  lz = (int)floor((z - f->z0) / f->dz);
  pz = (int)floor(((double)lz) / f->nz);
  lz -= pz * (int)f->nz;
  return lx + ly * f->nx + lz * f->nx * f->ny;
}

void orc_fdm_insert_energy(orc_fdm *f, double x, double y, double z, double E) { /* eph_fdm.h:172-179 */
  unsigned int index = (unsigned int)orc_fdm_index(f, x, y, z);
  double prescale = f->dV * f->dt;
  f->dT_e[index] += E / prescale;
}

double orc_fdm_get_T(const orc_fdm *f, double x, double y, double z) { /* eph_fdm.h:182-187 */
  return f->T_e[(unsigned int)orc_fdm_index(f, x, y, z)];
}

double orc_fdm_T_total(const orc_fdm *f) { /* eph_fdm.h:189-196 */
  double result = 0.;
  size_t i;
  for (i = 0; i < f->ntotal; ++i) result += f->T_e[i];
  result /= f->ntotal;
  return result;
}

void orc_fdm_solve(orc_fdm *f) { /* eph_fdm.h:267-400 (single rank: sync_before is an identity) */
  const size_t nx = f->nx, ny = f->ny, nz = f->nz, ntotal = f->ntotal;
  const double dx = f->dx, dy = f->dy, dz = f->dz, dt = f->dt;
  double *T_e = f->T_e, *dT_e = f->dT_e, *ddT_e = f->ddT_e, *C_e = f->C_e, *rho_e = f->rho_e, *kappa_e = f->kappa_e;
  const short *flag = f->flag;
  size_t i;
  unsigned int n, new_steps;
  double inner_dt = dt / f->steps;
  double dtdxdydz = inner_dt * (1.0 / dx / dx + 1.0 / dy / dy + 1.0 / dz / dz);
  double c_min, rho_min, kappa_max, r;

  for (i = 0; i < ntotal; ++i) /* :280-287 */
    if (f->T_dyn[i]) {
      C_e[i] = orc_spline_eval(f->C_e_T, 1. / f->dT, T_e[i]);
      kappa_e[i] = orc_spline_eval(f->kappa_e_T, 1. / f->dT, T_e[i]);
    }

  c_min = C_e[0]; rho_min = rho_e[0]; kappa_max = kappa_e[0]; /* seeded from cell 0 unconditionally, :290-292 */
  for (i = 1; i < ntotal; ++i)
    if (flag[i] != 0) {
      if (C_e[i] < c_min) c_min = C_e[i];
      if (rho_e[i] < rho_min) rho_min = rho_e[i];
      if (kappa_e[i] > kappa_max) kappa_max = kappa_e[i];
    }

  r = dtdxdydz / c_min / rho_min * kappa_max; /* :302 */
  new_steps = (unsigned int)f->steps;
  if (r > 0.4) { /* :308-313, truncating cast */
    unsigned int t;
    inner_dt = 0.4 * inner_dt / r;
    t = (unsigned int)(dt / inner_dt);
    new_steps = t > 1u ? t : 1u;
    inner_dt = dt / new_steps;
  }
  f->last_substeps = new_steps;

  for (n = 0; n < new_steps; ++n) {
    unsigned int ii, j, k;
    for (i = 0; i < ntotal; ++i) ddT_e[i] = 0.0;
    for (k = 0; k < nz; ++k)
      for (j = 0; j < ny; ++j)
        for (ii = 0; ii < nx; ++ii) { /* :319-367 */
          unsigned int q, p;
          unsigned int rr = ii + j * nx + k * nx * ny;
          if (flag[rr] == 2) continue;

          if (ii > 0) p = (ii - 1) + j * nx + k * nx * ny; else p = (nx - 1) + j * nx + k * nx * ny;
          if (ii < (nx - 1)) q = (ii + 1) + j * nx + k * nx * ny; else q = j * nx + k * nx * ny;
          if (flag[q] == 2) q = rr; else if (flag[p] == 2) p = rr;
          ddT_e[rr] += (kappa_e[q] - kappa_e[p]) * (T_e[q] - T_e[p]) / dx / dx / 4.0;
          ddT_e[rr] += kappa_e[rr] * ((T_e[q] + T_e[p] - 2.0 * T_e[rr]) / dx / dx);

          if (j > 0) p = ii + (j - 1) * nx + k * nx * ny; else p = ii + (ny - 1) * nx + k * nx * ny;
          if (j < (ny - 1)) q = ii + (j + 1) * nx + k * nx * ny; else q = ii + k * nx * ny;
          if (flag[q] == 2) q = rr; else if (flag[p] == 2) p = rr;
          ddT_e[rr] += (kappa_e[q] - kappa_e[p]) * (T_e[q] - T_e[p]) / dy / dy / 4.0;
          ddT_e[rr] += kappa_e[rr] * ((T_e[q] + T_e[p] - 2.0 * T_e[rr]) / dy / dy);

          if (k > 0) p = ii + j * nx + (k - 1) * nx * ny; else p = ii + j * nx + (nz - 1) * nx * ny;
          if (k < (nz - 1)) q = ii + j * nx + (k + 1) * nx * ny; else q = ii + j * nx;
          if (flag[q] == 2) q = rr; else if (flag[p] == 2) p = rr;
          ddT_e[rr] += (kappa_e[q] - kappa_e[p]) * (T_e[q] - T_e[p]) / dz / dz / 4.0;
          ddT_e[rr] += kappa_e[rr] * ((T_e[q] + T_e[p] - 2.0 * T_e[rr]) / dz / dz);
        }

    for (i = 0; i < ntotal; i++) { /* :371-395 */
      double prescaler = rho_e[i] * C_e[i];
      if (flag[i] == 1) {
        if (f->T_dyn[i] == 1) {
          double E_e = orc_linear_eval(f->dT, f->E_e_T, f->n_T, T_e[i]);
          E_e += (ddT_e[i] + dT_e[i] + f->S_e[i]) / rho_e[i] * inner_dt;
          T_e[i] = orc_linear_reverse(f->dT, f->E_e_T, f->n_T, E_e);
        } else {
          T_e[i] += (ddT_e[i] + dT_e[i] + f->S_e[i]) / prescaler * inner_dt;
        }
      }
      if (T_e[i] < 0.0) T_e[i] = 0.0;
    }
  }

  for (i = 0; i < ntotal; ++i) dT_e[i] = 0.0; /* sync_after, :484-491 */
}

int orc_fdm_save_temperature(const orc_fdm *f, const char *filename, int n) { /* eph_fdm.h:198-224 */
  char fn[1100];
  FILE *fd;
  size_t i, j, k;
  snprintf(fn, sizeof fn, "%s_%06d", filename, n);
  fd = fopen(fn, "w");
  if (!fd) return -1;
  fprintf(fd, "x y z Te\n");
  for (k = 0; k < f->nz; ++k)
    for (j = 0; j < f->ny; ++j)
      for (i = 0; i < f->nx; ++i) {
        unsigned int index = (unsigned int)(i + j * f->nx + k * f->nx * f->ny);
        double x = f->x0 + (int)i * f->dx, y = f->y0 + (int)j * f->dy, z = f->z0 + (int)k * f->dz;
        fprintf(fd, "%.6e %.6e %.6e %.6e\n", x, y, z, f->T_e[index]);
      }
  fclose(fd);
  return 0;
}

int orc_fdm_save_state(const orc_fdm *f, const char *filename) { /* eph_fdm.h:226-265 */
  FILE *fd = fopen(filename, "w");
  size_t i, j, k;
  if (!fd) return -1;
  fprintf(fd, "# A comment\n#\n#\n");
  fprintf(fd, "%ld %ld %ld %ld\n", (long)f->nx, (long)f->ny, (long)f->nz, (long)f->steps);
  fprintf(fd, "%.6e %.6e\n%.6e %.6e\n%.6e %.6e\n", f->x0, f->x1, f->y0, f->y1, f->z0, f->z1);
  fprintf(fd, "%s\n", f->parameter_filename);
  for (k = 0; k < f->nz; ++k)
    for (j = 0; j < f->ny; ++j)
      for (i = 0; i < f->nx; ++i) {
        unsigned int index = (unsigned int)(i + j * f->nx + k * f->nx * f->ny);
        fprintf(fd, "%d %d %d %.6e %.6e %.6e %.6e %.6e %d %d\n", (int)i, (int)j, (int)k, f->T_e[index], f->S_e[index],
                f->rho_e[index], f->C_e[index], f->kappa_e[index], f->flag[index], f->T_dyn[index]);
      }
  fclose(fd);
  return 0;
}

double *orc_fdm_field(orc_fdm *f, int which) {
  switch (which) {
    case 0: return f->T_e; case 1: return f->S_e; case 2: return f->rho_e;
    case 3: return f->C_e; case 4: return f->kappa_e; default: return f->dT_e;
  }
}
short *orc_fdm_flags(orc_fdm *f) { return f->flag; }
unsigned short *orc_fdm_tdyn(orc_fdm *f) { return f->T_dyn; }

/* =========================================================================
 * Counter-based Gaussian stream (our definition; the CUDA path implements the
 * same function).  xi_i(seed, step, tag): Philox4x32-10 with
 *   key = (seed_lo, seed_hi), counter = (tag_lo, tag_hi, step_lo, step_hi),
 * 32-bit uniforms u = (r + 0.5) / 2^32 and Box-Muller:
 *   xi0 = sqrt(-2 ln u0) cos(2 pi u1), xi1 = sqrt(-2 ln u0) sin(2 pi u1),
 *   xi2 = sqrt(-2 ln u2) cos(2 pi u3).
 * ========================================================================= */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  int r;
  for (r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_xi_stream(uint64_t seed, uint64_t step, long long n, const long long *tag, double *xi) {
  const double two_pi = 6.283185307179586476925286766559;
  const double scale = 1.0 / 4294967296.0;
  long long i;
  for (i = 0; i < n; ++i) {
    uint64_t t = (uint64_t)tag[i];
    uint32_t ctr[4] = {(uint32_t)t, (uint32_t)(t >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
    double u0, u1, u2, u3, m;
    orc_philox4x32_10(ctr, key, r);
    u0 = (r[0] + 0.5) * scale; u1 = (r[1] + 0.5) * scale; u2 = (r[2] + 0.5) * scale; u3 = (r[3] + 0.5) * scale;
    m = sqrt(-2.0 * log(u0));
    xi[3 * i + 0] = m * cos(two_pi * u1);
    xi[3 * i + 1] = m * sin(two_pi * u1);
    xi[3 * i + 2] = sqrt(-2.0 * log(u2)) * cos(two_pi * u3);
  }
}

/* =========================================================================
 * FixEPH hot path -- fix_eph.cpp
 * ========================================================================= */
orc_fix *orc_fix_new(int flags, int model, int groupbit, int ntypes, const int *type_map, double dt, double boltz,
                     double ftm2v, const orc_beta *beta, orc_fdm *fdm) {
  orc_fix *fx = (orc_fix *)calloc(1, sizeof(orc_fix));
  int t;
  fx->flags = flags; fx->model = model; fx->groupbit = groupbit; fx->ntypes = ntypes;
  for (t = 0; t < ntypes && t < 16; ++t) fx->type_map[t] = type_map[t];
  fx->boltz = boltz; fx->ftm2v = ftm2v; fx->beta = beta; fx->fdm = fdm;
  orc_fix_set_dt(fx, dt);
  return fx;
}

void orc_fix_free(orc_fix *fx) {
  if (!fx) return;
  free(fx->rho_i); free(fx->w_i); free(fx->xi_i); free(fx->f_EPH); free(fx->f_RNG); free(fx->array);
  free(fx->f_dis_i); free(fx->f_sto_i); free(fx);
}

void orc_fix_set_dt(orc_fix *fx, double dt) { /* fix_eph.cpp:909-916 */
  fx->dt = dt;
  fx->eta_factor = sqrt(2.0 * fx->boltz / dt);
  if (fx->coloured) fx->zeta_factor = 1.0 - exp(-dt / fx->tau0); /* fix_eph_coloured_exp.cpp:686 */
  if (fx->fdm) orc_fdm_set_dt(fx->fdm, dt);
}

static void fix_resize(orc_fix *fx, size_t n) { /* grow_arrays, fix_eph.cpp:918-938; new storage zeroed as in :229-238 */
  if (n <= fx->cap) return;
#define GROW(p, w) do { p = (double *)realloc(p, sizeof(double) * (w) * n); \
    memset(p + (w) * fx->cap, 0, sizeof(double) * (w) * (n - fx->cap)); } while (0)
  GROW(fx->rho_i, 1); GROW(fx->w_i, 3); GROW(fx->xi_i, 3); GROW(fx->f_EPH, 3); GROW(fx->f_RNG, 3); GROW(fx->array, 8);
  GROW(fx->f_dis_i, 3); GROW(fx->f_sto_i, 3);
#undef GROW
  fx->cap = n;
}

/* `fix eph/coloured/exp` (fix_eph_coloured_exp.cpp): the same path with an exponential memory kernel on both forces,
 * zeta_factor = 1 - exp(-dt / tau0) (:190, :686); tau0 <= 0 switches it off */
void orc_fix_set_colour(orc_fix *fx, double tau0) {
  fx->tau0 = tau0;
  fx->coloured = tau0 > 0.0;
  fx->zeta_factor = fx->coloured ? 1.0 - exp(-fx->dt / tau0) : 0.0;
}

static double dist_sq(const double *x, const double *y) { /* fix_eph.h:174-181 */
  double d0 = x[0] - y[0], d1 = x[1] - y[1], d2 = x[2] - y[2];
  return d0 * d0 + d1 * d1 + d2 * d2;
}
static double diff_sq(const double *x, const double *y, double *z) { /* fix_eph.h:184-191 */
  z[0] = x[0] - y[0]; z[1] = x[1] - y[1]; z[2] = x[2] - y[2];
  return z[0] * z[0] + z[1] * z[1] + z[2] * z[2];
}
static double dot3(const double *x, const double *y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }

/* Comm::forward_comm(Fix*) for one rank with periodic images: every ghost
 * receives its owner's value (pack/unpack, fix_eph.cpp:951-1009) */
static void ghost_fill(const orc_atoms *a, double *arr, int width) {
  int g, d;
  for (g = 0; g < a->nghost; ++g)
    for (d = 0; d < width; ++d) arr[(size_t)(a->nlocal + g) * width + d] = arr[(size_t)a->ghost_owner[g] * width + d];
}

void orc_calculate_environment(orc_fix *fx, const orc_atoms *a) { /* fix_eph.cpp:431-466 */
  const double rc2 = fx->beta->r_cutoff_sq;
  int i;
  fix_resize(fx, (size_t)a->nlocal + a->nghost);
  for (i = 0; i != a->nlocal; ++i) {
    fx->rho_i[i] = 0;
    if (a->mask[i] & fx->groupbit) {
      long long j;
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype = a->type[jj];
        double r_sq = dist_sq(a->x + 3 * (size_t)jj, a->x + 3 * (size_t)i);
        if (r_sq < rc2) fx->rho_i[i] += orc_beta_rho_r_sq(fx->beta, fx->type_map[jtype - 1], r_sq);
      }
    }
  }
}

void orc_force_prl(orc_fix *fx, const orc_atoms *a) { /* fix_eph.cpp:687-837 */
  const orc_beta *b = fx->beta;
  const double rc2 = b->r_cutoff_sq;
  const double *x = a->x, *v = a->v;
  double *rho_i = fx->rho_i, *w_i = fx->w_i, *xi_i = fx->xi_i;
  int i;
  if (fx->flags & ORC_FRICTION) {
    for (i = 0; i != a->nlocal; ++i) { /* w_i = W_ij^T v_j, :702-741 */
      long long j;
      double alpha_i;
      if (!(a->mask[i] & fx->groupbit)) continue;
      if (!(rho_i[i] > 0)) continue;
      alpha_i = orc_beta_alpha(b, fx->type_map[a->type[i] - 1], rho_i[i]);
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype = a->type[jj];
        double e_ij[3], e_r_sq = diff_sq(x + 3 * (size_t)jj, x + 3 * (size_t)i, e_ij);
        double v_rho_ji, prescaler, var1, var2, dvar;
        if (e_r_sq >= rc2) continue;
        v_rho_ji = orc_beta_rho_r_sq(b, fx->type_map[jtype - 1], e_r_sq);
        prescaler = alpha_i * v_rho_ji / (rho_i[i] * e_r_sq);
        var1 = prescaler * dot3(e_ij, v + 3 * (size_t)i);
        var2 = prescaler * dot3(e_ij, v + 3 * (size_t)jj);
        dvar = var1 - var2;
        w_i[3 * (size_t)i + 0] += dvar * e_ij[0];
        w_i[3 * (size_t)i + 1] += dvar * e_ij[1];
        w_i[3 * (size_t)i + 2] += dvar * e_ij[2];
      }
    }
    ghost_fill(a, w_i, 3); /* FixState::WI, :743-744 */
    for (i = 0; i != a->nlocal; ++i) { /* f_i = W_ij w_j, :748-787 */
      long long j;
      int itype = a->type[i];
      double alpha_i;
      if (!(a->mask[i] & fx->groupbit)) continue;
      if (!(rho_i[i] > 0)) continue;
      alpha_i = orc_beta_alpha(b, fx->type_map[itype - 1], rho_i[i]);
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype = a->type[jj];
        double e_ij[3], e_r_sq = diff_sq(x + 3 * (size_t)jj, x + 3 * (size_t)i, e_ij);
        double alpha_j, v_rho_ji, v_rho_ij, var1, var2, dvar;
        if (e_r_sq >= rc2 || !(rho_i[jj] > 0)) continue;
        alpha_j = orc_beta_alpha(b, fx->type_map[jtype - 1], rho_i[jj]);
        v_rho_ji = orc_beta_rho_r_sq(b, fx->type_map[jtype - 1], e_r_sq);
        var1 = alpha_i * v_rho_ji * dot3(e_ij, w_i + 3 * (size_t)i) / (rho_i[i] * e_r_sq);
        v_rho_ij = orc_beta_rho_r_sq(b, fx->type_map[itype - 1], e_r_sq);
        var2 = alpha_j * v_rho_ij * dot3(e_ij, w_i + 3 * (size_t)jj) / (rho_i[jj] * e_r_sq);
        dvar = var1 - var2;
        fx->f_EPH[3 * (size_t)i + 0] -= dvar * e_ij[0];
        fx->f_EPH[3 * (size_t)i + 1] -= dvar * e_ij[1];
        fx->f_EPH[3 * (size_t)i + 2] -= dvar * e_ij[2];
      }
      if (fx->coloured) { /* fix_eph_coloured_exp.cpp:563-569 */
        int d;
        for (d = 0; d < 3; ++d) {
          double *fd = fx->f_dis_i + 3 * (size_t)i + d;
          *fd = *fd * (1. - fx->zeta_factor) + fx->zeta_factor * fx->f_EPH[3 * (size_t)i + d];
          fx->f_EPH[3 * (size_t)i + d] = *fd;
        }
      }
    }
  }
  if (fx->flags & ORC_RANDOM) { /* :791-836 */
    for (i = 0; i != a->nlocal; i++) {
      long long j;
      int itype = a->type[i];
      double alpha_i, v_Te, var;
      if (!(a->mask[i] & fx->groupbit)) continue;
      if (!(rho_i[i] > 0)) continue;
      alpha_i = orc_beta_alpha(b, fx->type_map[itype - 1], rho_i[i]);
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype = a->type[jj];
        double e_ij[3], e_r_sq = diff_sq(x + 3 * (size_t)jj, x + 3 * (size_t)i, e_ij);
        double alpha_j, v_rho_ji, v_rho_ij, var1, var2, dvar;
        if ((e_r_sq >= rc2) || !(rho_i[jj] > 0)) continue;
        alpha_j = orc_beta_alpha(b, fx->type_map[jtype - 1], rho_i[jj]);
        v_rho_ji = orc_beta_rho_r_sq(b, fx->type_map[jtype - 1], e_r_sq);
        var1 = alpha_i * v_rho_ji * dot3(e_ij, xi_i + 3 * (size_t)i) / (rho_i[i] * e_r_sq);
        v_rho_ij = orc_beta_rho_r_sq(b, fx->type_map[itype - 1], e_r_sq);
        var2 = alpha_j * v_rho_ij * dot3(e_ij, xi_i + 3 * (size_t)jj) / (rho_i[jj] * e_r_sq);
        dvar = var1 - var2;
        fx->f_RNG[3 * (size_t)i + 0] += dvar * e_ij[0];
        fx->f_RNG[3 * (size_t)i + 1] += dvar * e_ij[1];
        fx->f_RNG[3 * (size_t)i + 2] += dvar * e_ij[2];
      }
      v_Te = orc_fdm_get_T(fx->fdm, x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2]);
      var = fx->eta_factor * sqrt(v_Te);
      fx->f_RNG[3 * (size_t)i + 0] *= var;
      fx->f_RNG[3 * (size_t)i + 1] *= var;
      fx->f_RNG[3 * (size_t)i + 2] *= var;
      if (fx->coloured) { /* fix_eph_coloured_exp.cpp:619-625 */
        int d;
        for (d = 0; d < 3; ++d) {
          double *fs = fx->f_sto_i + 3 * (size_t)i + d;
          *fs = *fs * (1. - fx->zeta_factor) + fx->zeta_factor * fx->f_RNG[3 * (size_t)i + d];
          fx->f_RNG[3 * (size_t)i + d] = *fs;
        }
      }
    }
  }
}

/* Random force shared by the two legacy models: f_RNG_i = eta_factor alpha(rho_i) sqrt(T_e(x_i)) xi_i for every group
 * atom (fix_eph.cpp:490-502 and :554-567 are the same loop). */
static void random_uncorrelated(orc_fix *fx, const orc_atoms *a) {
  const double *x = a->x;
  int i;
  for (i = 0; i != a->nlocal; ++i) {
    double v_Te, var;
    if (!(a->mask[i] & fx->groupbit)) continue;
    v_Te = orc_fdm_get_T(fx->fdm, x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2]);
    var = fx->eta_factor * orc_beta_alpha(fx->beta, fx->type_map[a->type[i] - 1], fx->rho_i[i]) * sqrt(v_Te);
    fx->f_RNG[3 * (size_t)i + 0] = var * fx->xi_i[3 * (size_t)i + 0];
    fx->f_RNG[3 * (size_t)i + 1] = var * fx->xi_i[3 * (size_t)i + 1];
    fx->f_RNG[3 * (size_t)i + 2] = var * fx->xi_i[3 * (size_t)i + 2];
  }
}

void orc_force_ttm(orc_fix *fx, const orc_atoms *a) { /* fix_eph.cpp:468-503 */
  const double *v = a->v;
  int i;
  if (fx->flags & ORC_FRICTION)
    for (i = 0; i != a->nlocal; ++i) {
      double var;
      if (!(a->mask[i] & fx->groupbit)) continue;
      var = -orc_beta_beta(fx->beta, fx->type_map[a->type[i] - 1], fx->rho_i[i]);
      fx->f_EPH[3 * (size_t)i + 0] = var * v[3 * (size_t)i + 0];
      fx->f_EPH[3 * (size_t)i + 1] = var * v[3 * (size_t)i + 1];
      fx->f_EPH[3 * (size_t)i + 2] = var * v[3 * (size_t)i + 2];
    }
  if (fx->flags & ORC_RANDOM) random_uncorrelated(fx, a);
}

void orc_force_prb(orc_fix *fx, const orc_atoms *a) { /* fix_eph.cpp:505-568 */
  const orc_beta *b = fx->beta;
  const double rc2 = b->r_cutoff_sq;
  const double *x = a->x, *v = a->v;
  int i;
  if (fx->flags & ORC_FRICTION)
    for (i = 0; i != a->nlocal; ++i) {
      long long j;
      double *f = fx->f_EPH + 3 * (size_t)i, var;
      if (!(a->mask[i] & fx->groupbit)) continue;
      if (!(fx->rho_i[i] > 0)) continue;
      f[0] = v[3 * (size_t)i + 0]; f[1] = v[3 * (size_t)i + 1]; f[2] = v[3 * (size_t)i + 2];
      for (j = a->offsets[i]; j != a->offsets[i + 1]; ++j) {
        int jj = a->neigh[j] & NEIGHMASK;
        int jtype = a->type[jj];
        double r_sq = dist_sq(x + 3 * (size_t)jj, x + 3 * (size_t)i);
        if (r_sq < rc2) { /* the element index is jtype - 1, NOT type_map[jtype - 1] (:530) */
          var = orc_beta_rho_r(b, jtype - 1, sqrt(r_sq)) / fx->rho_i[i];
          f[0] -= var * v[3 * (size_t)jj + 0];
          f[1] -= var * v[3 * (size_t)jj + 1];
          f[2] -= var * v[3 * (size_t)jj + 2];
        }
      }
      var = orc_beta_beta(b, fx->type_map[a->type[i] - 1], fx->rho_i[i]);
      f[0] *= var; f[1] *= var; f[2] *= var;
    }
  if (fx->flags & ORC_RANDOM) random_uncorrelated(fx, a);
}

void orc_post_force(orc_fix *fx, const orc_atoms *a, const double *xi) { /* fix_eph.cpp:841-907 */
  size_t nl = (size_t)a->nlocal, i;
  fix_resize(fx, nl + a->nghost);
  memset(fx->w_i, 0, sizeof(double) * 3 * nl); /* :848-851 */
  memset(fx->xi_i, 0, sizeof(double) * 3 * nl);
  memset(fx->f_EPH, 0, sizeof(double) * 3 * nl);
  memset(fx->f_RNG, 0, sizeof(double) * 3 * nl);
  if (fx->flags & ORC_RANDOM) { /* :854-865 */
    for (i = 0; i < nl; ++i)
      if (a->mask[i] & fx->groupbit) {
        fx->xi_i[3 * i + 0] = xi[3 * i + 0];
        fx->xi_i[3 * i + 1] = xi[3 * i + 1];
        fx->xi_i[3 * i + 2] = xi[3 * i + 2];
      }
    ghost_fill(a, fx->xi_i, 3);
  }
  orc_calculate_environment(fx, a); /* :868 */
  ghost_fill(a, fx->rho_i, 1);      /* :870-871 */
  if (fx->model == 1) orc_force_ttm(fx, a);       /* switch(eph_model), :877-890; PRLCM (3) is not restated: the */
  else if (fx->model == 2) orc_force_prb(fx, a);  /* reference indexes its table with jtype - i there (:601)      */
  else if (fx->model == 4) orc_force_prl(fx, a);
  if ((fx->flags & ORC_FRICTION) && !(fx->flags & ORC_NOFRICTION)) /* :892-898 */
    for (i = 0; i < 3 * nl; ++i) a->f[i] += fx->f_EPH[i];
  if ((fx->flags & ORC_RANDOM) && !(fx->flags & ORC_NORANDOM)) /* :900-906 */
    for (i = 0; i < 3 * nl; ++i) a->f[i] += fx->f_RNG[i];
}

void orc_end_of_step(orc_fix *fx, const orc_atoms *a) { /* fix_eph.cpp:350-429 */
  const double *x = a->x, *v = a->v;
  double E_local = 0.0;
  size_t nl = (size_t)a->nlocal, i;
  fix_resize(fx, nl + a->nghost);
  if (fx->flags & ORC_FRICTION)
    for (i = 0; i < nl; ++i)
      if (a->mask[i] & fx->groupbit) {
        double dE_i = 0.0;
        dE_i -= fx->f_EPH[3 * i + 0] * v[3 * i + 0] * fx->dt;
        dE_i -= fx->f_EPH[3 * i + 1] * v[3 * i + 1] * fx->dt;
        dE_i -= fx->f_EPH[3 * i + 2] * v[3 * i + 2] * fx->dt;
        orc_fdm_insert_energy(fx->fdm, x[3 * i], x[3 * i + 1], x[3 * i + 2], dE_i);
        E_local += dE_i;
      }
  if (fx->flags & ORC_RANDOM)
    for (i = 0; i < nl; ++i)
      if (a->mask[i] & fx->groupbit) {
        double dE_i = 0.0;
        dE_i -= fx->f_RNG[3 * i + 0] * v[3 * i + 0] * fx->dt;
        dE_i -= fx->f_RNG[3 * i + 1] * v[3 * i + 1] * fx->dt;
        dE_i -= fx->f_RNG[3 * i + 2] * v[3 * i + 2] * fx->dt;
        orc_fdm_insert_energy(fx->fdm, x[3 * i], x[3 * i + 1], x[3 * i + 2], dE_i);
        E_local += dE_i;
      }
  if (fx->flags & ORC_FDM) orc_fdm_solve(fx->fdm);
  fx->Ee += E_local;
  for (i = 0; i < nl; ++i) { /* :406-428 */
    double *row = fx->array + 8 * i;
    if (a->mask[i] & fx->groupbit) {
      row[0] = fx->rho_i[i];
      row[1] = orc_beta_beta(fx->beta, fx->type_map[a->type[i] - 1], fx->rho_i[i]);
      row[2] = fx->f_EPH[3 * i + 0]; row[3] = fx->f_EPH[3 * i + 1]; row[4] = fx->f_EPH[3 * i + 2];
      row[5] = fx->f_RNG[3 * i + 0]; row[6] = fx->f_RNG[3 * i + 1]; row[7] = fx->f_RNG[3 * i + 2];
    } else {
      memset(row, 0, sizeof(double) * 8);
    }
  }
}

void orc_initial_integrate(orc_fix *fx, const orc_atoms *a, const double *mass) { /* fix_eph.cpp:305-328 */
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

void orc_final_integrate(orc_fix *fx, const orc_atoms *a, const double *mass) { /* fix_eph.cpp:330-348 */
  double dtf = 0.5 * fx->dt * fx->ftm2v;
  int i, d;
  if (fx->flags & ORC_NOINT) return;
  for (i = 0; i < a->nlocal; ++i)
    if (a->mask[i] & fx->groupbit) {
      double dtfm = dtf / mass[a->type[i]];
      for (d = 0; d < 3; ++d) a->v[3 * (size_t)i + d] += dtfm * a->f[3 * (size_t)i + d];
    }
}

void orc_atoms_fill(orc_atoms *a, int nlocal, int nghost, double *x, double *v, double *f, const int *type,
                    const int *mask, const int *ghost_owner, const long long *offsets, const int *neigh) {
  a->nlocal = nlocal; a->nghost = nghost; a->x = x; a->v = v; a->f = f; a->type = type; a->mask = mask;
  a->ghost_owner = ghost_owner; a->offsets = offsets; a->neigh = neigh;
}
size_t orc_sizeof_atoms(void) { return sizeof(orc_atoms); }
double *orc_fix_ptr(orc_fix *fx, int which) {
  switch (which) {
    case 0: return fx->rho_i; case 1: return fx->w_i; case 2: return fx->xi_i;
    case 3: return fx->f_EPH; case 4: return fx->f_RNG; case 6: return fx->f_dis_i; case 7: return fx->f_sto_i;
    default: return fx->array;
  }
}
double orc_fix_Ee(const orc_fix *fx) { return fx->Ee; }
