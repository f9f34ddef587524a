/* TEST INFRASTRUCTURE -- the parity oracle, not the product.
 *
 * A plain-C, single-threaded restatement of the per-timestep `fix eph` hot
 * path of LLNL/USER-EPH.  Every function cites the
 * reference file:line it follows (paths relative to the reference checkout).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product (user-eph_b200/) never does.
 *
 * Pinning: tests/test_oracle_vs_reference.py checks this restatement
 * bit-for-bit against the UNMODIFIED reference compiled into
 * oracle/_ref/libeph_ref.so (tables, rho, w, f_EPH, f_RNG, FDM grids), and
 * tests/golden/ holds vectors generated from that compiled reference.
 */
#ifndef EPH_ORACLE_H
#define EPH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- EPH_Spline (eph_spline.h) ---- */
void orc_spline_build(double dx, const double *y, size_t n, double *coeff /* [n][4] = a,b,c,d */);
double orc_spline_eval(const double *coeff, double inv_dx, double x);

/* ---- EPH_Linear (eph_linear.h) ---- */
double orc_linear_eval(double dx, const double *y, size_t n, double x);
double orc_linear_reverse(double dx, const double *y, size_t n, double yv);

/* ---- EPH_Beta (eph_beta.h) ---- */
typedef struct orc_beta {
  int n_elements;
  size_t n_rho, n_beta;
  double dr, dr_sq, drho;
  double r_cutoff, r_cutoff_sq, rho_cutoff;
  double inv_dr, inv_dr_sq, inv_drho;
  char names[16][16];
  int number[16];
  double *rho_r;    /* [n_elements][n_rho][4]  */
  double *rho_r_sq; /* [n_elements][n_rho][4]  */
  double *alpha;    /* [n_elements][n_beta][4] */
  double *beta;     /* [n_elements][n_beta][4] */
} orc_beta;

orc_beta *orc_beta_load(const char *file);
/* same tables from in-memory knots: rho_knots[n_el][n_rho], beta_knots[n_el][n_beta] */
orc_beta *orc_beta_from_knots(int n_elements, size_t n_rho, double dr, size_t n_beta, double drho, double r_cutoff,
                              const double *rho_knots, const double *beta_knots);
void orc_beta_free(orc_beta *b);
double orc_beta_rho_r_sq(const orc_beta *b, int e, double r_sq);
double orc_beta_rho_r(const orc_beta *b, int e, double r);
double orc_beta_alpha(const orc_beta *b, int e, double rho);
double orc_beta_beta(const orc_beta *b, int e, double rho);
/* plain accessors for ctypes */
void orc_beta_info(const orc_beta *b, long long *dims /*3*/, double *scal /*6*/);
const double *orc_beta_table(const orc_beta *b, int kind /*0 rho(r) 1 rho(r^2) 2 alpha 3 beta*/, int e);

/* ---- EPH_FDM (eph_fdm.h) ---- */
typedef struct orc_fdm {
  size_t nx, ny, nz, ntotal, steps;
  double x0, x1, y0, y1, z0, z1, dx, dy, dz, dV, dt;
  double *T_e, *dT_e, *ddT_e, *C_e, *rho_e, *kappa_e, *S_e;
  short *flag;
  unsigned short *T_dyn;
  /* temperature dependent parameters (parameter file) */
  size_t n_T;
  double dT;
  double *C_e_T;     /* spline coeff [n_T][4] */
  double *kappa_e_T; /* spline coeff [n_T][4] */
  double *E_e_T;     /* linear table y [n_T] */
  char parameter_filename[1024];
  unsigned int last_substeps;
} orc_fdm;

orc_fdm *orc_fdm_new(size_t nx, size_t ny, size_t nz, const double *box /*x0 x1 y0 y1 z0 z1*/, double T_e, double C_e,
                     double rho_e, double kappa_e);
orc_fdm *orc_fdm_from_file(const char *file);
void orc_fdm_free(orc_fdm *f);
void orc_fdm_set_dt(orc_fdm *f, double dt);
size_t orc_fdm_index(const orc_fdm *f, double x, double y, double z);
void orc_fdm_insert_energy(orc_fdm *f, double x, double y, double z, double E);
double orc_fdm_get_T(const orc_fdm *f, double x, double y, double z);
double orc_fdm_T_total(const orc_fdm *f);
void orc_fdm_solve(orc_fdm *f);
int orc_fdm_save_temperature(const orc_fdm *f, const char *filename, int n);
int orc_fdm_save_state(const orc_fdm *f, const char *filename);
double *orc_fdm_field(orc_fdm *f, int which /*0 T_e 1 S_e 2 rho_e 3 C_e 4 kappa_e 5 dT_e*/);
short *orc_fdm_flags(orc_fdm *f);
unsigned short *orc_fdm_tdyn(orc_fdm *f);

/* ---- counter-based Gaussian stream shared with the CUDA path (ours, not the
 *      reference's RanMars, which is third-party and unpinned) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_xi_stream(uint64_t seed, uint64_t step, long long n, const long long *tag, double *xi /*[n][3]*/);

/* ---- FixEPH hot path (fix_eph.cpp), model 4 (PRL) ---- */
enum { ORC_FRICTION = 1, ORC_RANDOM = 2, ORC_FDM = 4, ORC_NOINT = 8, ORC_NOFRICTION = 16, ORC_NORANDOM = 32 };

typedef struct orc_fix {
  int flags, model, groupbit, ntypes;
  int type_map[16];
  double dt, boltz, ftm2v, eta_factor;
  const orc_beta *beta;
  orc_fdm *fdm;
  double Ee;
  /* per-atom state, sized by orc_fix_resize */
  size_t cap;
  double *rho_i, *w_i, *xi_i, *f_EPH, *f_RNG, *array;
  /* fix eph/coloured/exp (fix_eph_coloured_exp.cpp): exponential memory kernel on both forces */
  int coloured;
  double tau0, zeta_factor;
  double *f_dis_i, *f_sto_i; /* [n][3] filtered friction / random force, carried from step to step */
} orc_fix;

orc_fix *orc_fix_new(int flags, int model, int groupbit, int ntypes, const int *type_map, double dt, double boltz,
                     double ftm2v, const orc_beta *beta, orc_fdm *fdm);
void orc_fix_free(orc_fix *fx);
void orc_fix_set_dt(orc_fix *fx, double dt);
void orc_fix_set_colour(orc_fix *fx, double tau0);

typedef struct orc_atoms {
  int nlocal, nghost;
  double *x, *v, *f; /* [nlocal+nghost][3] */
  const int *type, *mask;
  const int *ghost_owner;      /* [nghost] local index of the owner */
  const long long *offsets;    /* CSR [nlocal+1] */
  const int *neigh;            /* raw LAMMPS entries (NEIGHMASK applied inside) */
} orc_atoms;

void orc_calculate_environment(orc_fix *fx, const orc_atoms *a);
void orc_force_prl(orc_fix *fx, const orc_atoms *a);
void orc_force_ttm(orc_fix *fx, const orc_atoms *a);   /* model 1, fix_eph.cpp:468-503 */
void orc_force_prb(orc_fix *fx, const orc_atoms *a);   /* model 2, fix_eph.cpp:505-568 */
/* xi: [nlocal][3] Gaussians for every local atom (only group atoms are used) or NULL when RANDOM is off */
void orc_post_force(orc_fix *fx, const orc_atoms *a, const double *xi);
void orc_end_of_step(orc_fix *fx, const orc_atoms *a);
void orc_initial_integrate(orc_fix *fx, const orc_atoms *a, const double *mass_by_type /*1-based*/);
void orc_final_integrate(orc_fix *fx, const orc_atoms *a, const double *mass_by_type);
/* flat-argument wrappers for ctypes */
void orc_atoms_fill(orc_atoms *a, int nlocal, int nghost, double *x, double *v, double *f, const int *type,
                    const int *mask, const int *ghost_owner, const long long *offsets, const int *neigh);
size_t orc_sizeof_atoms(void);
double *orc_fix_ptr(orc_fix *fx, int which /*0 rho 1 w 2 xi 3 f_EPH 4 f_RNG 5 array 6 f_dis 7 f_sto*/);
double orc_fix_Ee(const orc_fix *fx);

/* ---- `fix eph/atomic`: EPH_kappa (eph_kappa.h) and FixEPHAtomic (fix_eph_atomic.cpp);
 *      restated in eph_oracle_atomic.c ---- */
typedef struct orc_kappa {
  int n_elements, n_pairs;
  size_t n_r, n_T;
  double r_cutoff, r_cutoff_sq, T_max, dT, inv_dr, inv_dr_sq;
  char names[16][16];
  int number[16];
  double *rho_r;    /* [n_elements][n_r][4] */
  double *rho_r_sq; /* [n_elements][n_r][4] */
  double *E_T;      /* [n_elements][n_T] running sum of C(T) dT (EPH_Linear y) */
  double *K_T;      /* [n_pairs][n_T] (EPH_Linear y) */
} orc_kappa;

orc_kappa *orc_kappa_load(const char *file);
void orc_kappa_free(orc_kappa *k);
void orc_kappa_info(const orc_kappa *k, long long *dims /*4*/, double *scal /*5*/);
const double *orc_kappa_table(const orc_kappa *k, int kind /*0 rho(r) 1 rho(r^2) 2 E(T) 3 K(T)*/, int e);

typedef struct orc_afix {
  int flags, groupbit, ntypes, inner_loops;
  int type_map_beta[16], type_map_kappa[16];
  double dt, boltz, ftm2v, eta_factor;
  const orc_beta *beta;
  const orc_kappa *kappa;
  double Ee, Te;
  size_t cap;
  double *rho_i, *w_i, *xi_i, *f_EPH, *f_RNG, *array /*[n][12]*/;
  double *rho_a_i, *E_a_i /*[n][2]*/, *dE_a_i, *T_a_i;
} orc_afix;

/* `a` supplies nlocal/nghost/type/mask for the constructor's per-atom initialisation (fix_eph_atomic.cpp:212-257) */
orc_afix *orc_afix_new(int flags, int groupbit, int ntypes, const int *type_map_beta, const int *type_map_kappa, double dt,
                       double boltz, double ftm2v, int inner_loops, double T_init, const orc_beta *beta,
                       const orc_kappa *kappa, const orc_atoms *a);
void orc_afix_free(orc_afix *fx);
void orc_afix_set_dt(orc_afix *fx, double dt);
void orc_atomic_calculate_environment(orc_afix *fx, const orc_atoms *a);
void orc_atomic_force_prl(orc_afix *fx, const orc_atoms *a);
void orc_atomic_heat_solve(orc_afix *fx, const orc_atoms *a);
void orc_atomic_post_force(orc_afix *fx, const orc_atoms *a, const double *xi);
void orc_atomic_end_of_step(orc_afix *fx, const orc_atoms *a);
void orc_atomic_initial_integrate(orc_afix *fx, const orc_atoms *a, const double *mass_by_type);
void orc_atomic_final_integrate(orc_afix *fx, const orc_atoms *a, const double *mass_by_type);
/* which: 0 rho 1 w 2 xi 3 f_EPH 4 f_RNG 5 rho_a 6 E_a[n][2] 7 dE_a 8 T_a 9 array[n][12] */
double *orc_afix_ptr(orc_afix *fx, int which);
double orc_afix_Ee(const orc_afix *fx);
double orc_afix_Te(const orc_afix *fx);

#ifdef __cplusplus
}
#endif
#endif
