/* eph_b200_atomic -- C ABI of the B200-native `fix eph/atomic` path.
 *
 * `fix eph/atomic` is the reference's variant of `fix eph` that carries the electronic energy per ATOM instead of on
 * an FDM grid: the same density / friction / random sweeps over the full neighbour list, an energy ledger dE_i fed
 * pair by pair, and a heat-diffusion step between neighbouring atoms (fix_eph_atomic.cpp, eph_kappa.h).  This header
 * is its drop-in boundary, with the conventions of eph_b200.h: plain C, opaque handle, EPH_B200_OK or a negative
 * eph_b200_status, one CUDA stream, LAMMPS array layouts, host or device pointers (`memspace`), NO CPU fallback.
 *
 * Ghost atoms take their owner's values either through the ghost_owner map inside the engine (one rank per box: every
 * ghost is a periodic image of one of the rank's own atoms) or through the caller's transport in a phase-split step
 * (set_comm_mode; what Comm::forward_comm(Fix*) realises through pack/unpack_forward_comm, fix_eph_atomic.cpp:849-927).
 * Every entry point names the reference lines it replaces.
 */
#ifndef EPH_B200_ATOMIC_H
#define EPH_B200_ATOMIC_H

#include "eph_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eph_b200_atomic_handle eph_b200_atomic_handle;

/* The scalar part of FixEPHAtomic::FixEPHAtomic (fix_eph_atomic.cpp:58-258).  flags: FixEPHAtomic::Flag
 * (fix_eph_atomic.h:44-51; the bit EPH_B200_FDM = 4 is HEAT here). */
typedef struct eph_b200_atomic_config {
  int device;
  int ntypes;
  const int *type_map_beta;  /* [ntypes] element index in the .beta file per LAMMPS type  (:189-196) */
  const int *type_map_kappa; /* [ntypes] element index in the .kappa file per LAMMPS type (:198-203) */
  int groupbit;
  int flags;                 /* arg[4] */
  unsigned long long seed;   /* arg[3]; keys the counter-based Gaussian stream */
  int inner_loops;           /* arg[7]; < 1 means one loop (:159-163, :695-699) */
  void *stream;              /* cudaStream_t or NULL */
} eph_b200_atomic_config;

int eph_b200_atomic_create(const eph_b200_atomic_config *cfg, eph_b200_atomic_handle **out);
int eph_b200_atomic_destroy(eph_b200_atomic_handle *h);
const char *eph_b200_atomic_last_error(const eph_b200_atomic_handle *h);
const char *eph_b200_atomic_create_error(void);
long long eph_b200_atomic_launch_count(const eph_b200_atomic_handle *h);
int eph_b200_atomic_synchronize(eph_b200_atomic_handle *h);

/* EPH_Beta's spline set, as eph_b200_set_tables (eph_beta.h:96-125, :164-198). */
int eph_b200_atomic_set_beta_tables(eph_b200_atomic_handle *h, int n_elements, int n_rho, double inv_dr_sq,
                                    const double *coeff_rho_r_sq, int n_beta, double inv_drho, const double *coeff_alpha,
                                    const double *coeff_beta, double r_cutoff_sq, double rho_cutoff);
/* EPH_kappa's tables (eph_kappa.h:53-151): locality splines rho_a(r^2) [n_elements][n_r][4] in absolute x, the
 * running-sum table E(T) [n_elements][n_T] and the conductivity table K(T) [n_pairs][n_T] (both EPH_Linear knots
 * spaced dT).  The reference indexes K(T) by ELEMENT (fix_eph_atomic.cpp:731, :747), so n_pairs >= n_elements is
 * required (its own n_pairs formula, eph_kappa.h:69, gives 1 for two elements: such files are rejected here, the
 * reference reads out of bounds with them). */
int eph_b200_atomic_set_kappa_tables(eph_b200_atomic_handle *h, int n_elements, int n_pairs, int n_r, double inv_dr_sq,
                                     const double *coeff_rho_r_sq, double r_cutoff_sq, int n_T, double dT,
                                     const double *E_T, const double *K_T);
/* FixEPHAtomic::reset_dt (fix_eph_atomic.cpp:809-814) */
int eph_b200_atomic_set_dt(eph_b200_atomic_handle *h, double dt, double boltz);

/* atom->type, mask, tag for nlocal+nghost atoms and the ghost->owner map (all ghosts need an owner >= 0). */
int eph_b200_atomic_set_atoms(eph_b200_atomic_handle *h, int nlocal, int nghost, const int *type, const int *mask,
                              const int64_t *tag, const int *ghost_owner, int memspace);
/* list->numneigh / firstneigh as CSR (fix_eph_atomic.cpp:443-444); call when neighbor->ago == 0 */
int eph_b200_atomic_set_neighbors_csr(eph_b200_atomic_handle *h, int nlocal, const int64_t *offsets, const int *neigh,
                                      int memspace);

/* Per-atom electronic energy E_a_i[.][0].  init: E_i = E(T_init) of the atom's element for group atoms, 0 otherwise
 * (constructor, fix_eph_atomic.cpp:212-221).  set/get move the local atoms' values (LAMMPS re-orders and migrates
 * atoms between steps: the host fix keeps E_a_i in its own per-atom array, as the reference does through
 * copy_arrays / pack_exchange / unpack_exchange, :939-955, and re-registers it when the atoms were re-ordered). */
int eph_b200_atomic_init_energy(eph_b200_atomic_handle *h, double T_init);
int eph_b200_atomic_set_energy(eph_b200_atomic_handle *h, const double *E, int memspace);
int eph_b200_atomic_get_energy(eph_b200_atomic_handle *h, double *E, int memspace);

/* FixEPHAtomic::post_force (fix_eph_atomic.cpp:789-857): xi, calculate_environment (:437-490; rho_i and the locality
 * density rho_a_i, neighbours outside the fix group skipped), the ghost fills, force_prl (:492-679) with the energy
 * ledger dE_a_i, and f += f_EPH (+ f_RNG) for GROUP atoms.  x, v: [nlocal+nghost][3]; f: [nlocal][3]. */
int eph_b200_atomic_post_force(eph_b200_atomic_handle *h, const double *x, const double *v, double *f,
                               const double *xi_inject, long long ntimestep, int memspace);
/* FixEPHAtomic::end_of_step (fix_eph_atomic.cpp:361-399): heat_solve (:681-787) when flag HEAT is set, then the
 * group's energy sum (f_ID[1]) and mean temperature (f_ID[2]); T_a_i is refreshed.  Uses the positions of the last
 * post_force.  Ee / Te may be NULL (no host sync then). */
int eph_b200_atomic_end_of_step(eph_b200_atomic_handle *h, double *Ee, double *Te);
/* only the sums (constructor, fix_eph_atomic.cpp:223-253) */
int eph_b200_atomic_summary(eph_b200_atomic_handle *h, double *Ee, double *Te);

/* Ghost transport by the caller (LAMMPS' Comm::forward_comm(Fix*) with host buffers, i.e. several ranks per box): after
 * set_comm_mode(h, 1) set_atoms needs no owner map and a step is driven in phases, with the reference's forward comms
 * (fix_eph_atomic.cpp:803-825, :549-550, :720-721) issued by the caller through pack_forward / unpack_forward
 * (FixEPHAtomic::pack_forward_comm / unpack_forward_comm, :849-927; state 1 RHO {rho, rho_a}, 2 XI, 3 WI, 4 EI):
 *   post_force_begin;  comm EI, XI (flag RANDOM), RHO;  post_force_mid;  comm WI (flag FRICTION);  post_force_end;
 *   heat_loops() x { heat_begin;  comm EI;  heat_end };  summary (this rank's energy sum and mean temperature).
 * pack_forward returns the number of doubles written or a negative status. */
int eph_b200_atomic_set_comm_mode(eph_b200_atomic_handle *h, int external);
int eph_b200_atomic_post_force_begin(eph_b200_atomic_handle *h, const double *x, const double *v, const double *xi_inject,
                                     long long ntimestep, int memspace);
int eph_b200_atomic_post_force_mid(eph_b200_atomic_handle *h);
int eph_b200_atomic_post_force_end(eph_b200_atomic_handle *h, double *f, int memspace);
int eph_b200_atomic_heat_loops(const eph_b200_atomic_handle *h);
int eph_b200_atomic_heat_begin(eph_b200_atomic_handle *h);
int eph_b200_atomic_heat_end(eph_b200_atomic_handle *h);
int eph_b200_atomic_pack_forward(eph_b200_atomic_handle *h, int state, int n, const int *list, double *buf);
int eph_b200_atomic_unpack_forward(eph_b200_atomic_handle *h, int state, int n, int first, const double *buf);

/* FixEPHAtomic::populate_array (fix_eph_atomic.cpp:401-435): [nlocal][12] = rho, beta(rho), f_EPH xyz, f_RNG xyz,
 * rho_a, E, dE, T (zeros for atoms outside the group) */
int eph_b200_atomic_get_peratom(eph_b200_atomic_handle *h, double *array12, int memspace);
/* probes for parity tests (host pointers).  which: 0 rho[nt] 1 w[nl][3] 2 xi[nl][3] 3 f_EPH[nl][3] 4 f_RNG[nl][3]
 * 5 rho_a[nt] 6 E[nt] 7 dE[nl] 8 T[nl] */
int eph_b200_atomic_get_probe(eph_b200_atomic_handle *h, int which, double *out);

#ifdef __cplusplus
}
#endif
#endif /* EPH_B200_ATOMIC_H */
