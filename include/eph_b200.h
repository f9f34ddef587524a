/* eph_b200 -- C ABI of the B200-native `fix eph` hot path.
 *
 * This is the drop-in boundary: a LAMMPS-side `FixEPH`-compatible host class
 * (user-eph_b200/fix/fix_eph_b200.cpp) keeps the reference's argument syntax,
 * file formats and Fix hooks and forwards the per-timestep work to these entry
 * points.  Each entry point names the reference interface it replaces
 * (paths relative to the LLNL/USER-EPH checkout).
 *
 * Conventions
 *  - plain C, opaque handle, no exceptions across the boundary;
 *  - every call returns EPH_B200_OK (0) or a negative error code, and
 *    eph_b200_last_error(h) gives the message (the host fix forwards it to
 *    LAMMPS' error->all(), mirroring fix_eph.cpp:64,140,175,182,196);
 *  - all calls come from the rank's single host thread (LAMMPS calls fix hooks
 *    from the main thread); work is enqueued on ONE CUDA stream (config.stream
 *    or a library-owned one);
 *  - per-atom arrays use LAMMPS layout: x, v, f are [n][3] row-major doubles
 *    (`&atom->x[0][0]`), type/mask int32, tag int64; ghosts follow locals;
 *  - pointer arguments are HOST pointers unless the `memspace` argument says
 *    EPH_B200_DEVICE, in which case they are device pointers of the same
 *    layout (GPU-resident LAMMPS: KOKKOS / GPU package) and no copy is made;
 *  - there is NO CPU fallback: without a CUDA device create() fails.
 */
#ifndef EPH_B200_H
#define EPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPH_B200_VERSION 100

enum eph_b200_status {
  EPH_B200_OK = 0,
  EPH_B200_ERR_ARG = -1,      /* bad argument / call order                    */
  EPH_B200_ERR_CUDA = -2,     /* CUDA runtime error (message has the detail)  */
  EPH_B200_ERR_NODEVICE = -3, /* no usable sm_100 device                       */
  EPH_B200_ERR_MODEL = -4,    /* friction model not available on the device   */
  EPH_B200_ERR_STATE = -5,    /* numerical guard tripped (see status word)    */
  EPH_B200_ERR_COMM = -6      /* NCCL missing or an NCCL call failed           */
};

enum eph_b200_memspace { EPH_B200_HOST = 0, EPH_B200_DEVICE = 1 };

/* eph_flag bits -- FixEPH::Flag, fix_eph.h:43-50 */
enum eph_b200_flag {
  EPH_B200_FRICTION = 0x01,
  EPH_B200_RANDOM = 0x02,
  EPH_B200_FDM = 0x04,
  EPH_B200_NOINT = 0x08,
  EPH_B200_NOFRICTION = 0x10,
  EPH_B200_NORANDOM = 0x20
};

/* eph_model -- FixEPH::Model, fix_eph.h:53-60.  PRL (4) is the model every BASELINE configuration uses; TTM (1,
 * fix_eph.cpp:468-503) and PRB (2, :505-568) are the reference's uncorrelated legacy models (PRB additionally needs
 * eph_b200_set_rho_r_table).  PRLCM (3) is rejected: the reference reads out of bounds there (fix_eph.cpp:601). */
enum eph_b200_model { EPH_B200_MODEL_NONE = 0, EPH_B200_MODEL_TTM = 1, EPH_B200_MODEL_PRB = 2, EPH_B200_MODEL_PRL = 4 };

/* forward-comm payloads -- FixEPH::FixState, fix_eph.h:35-40 */
enum eph_b200_state { EPH_B200_STATE_NONE = 0, EPH_B200_STATE_RHO = 1, EPH_B200_STATE_XI = 2, EPH_B200_STATE_WI = 3 };

typedef struct eph_b200_handle eph_b200_handle;

/* Replaces the scalar part of FixEPH::FixEPH (fix_eph.cpp:61-241): what the
 * constructor parses from the `fix ... eph` command line and LAMMPS state. */
typedef struct eph_b200_config {
  int device;             /* CUDA device ordinal of this rank                          */
  int ntypes;             /* atom->ntypes                                              */
  const int *type_map;    /* [ntypes] element index in the .beta file per LAMMPS type  */
  int groupbit;           /* Fix::groupbit                                             */
  int flags;              /* eph_flag, arg[4]                                          */
  int model;              /* eph_model, arg[5]                                         */
  unsigned long long seed;/* arg[3]; keys the counter-based Gaussian stream            */
  int rank, nranks;       /* comm->me, comm->nprocs                                    */
  void *stream;           /* cudaStream_t to enqueue on, or NULL for a private stream  */
} eph_b200_config;

int eph_b200_version(void);
/* number of CUDA devices this process sees (the host classes map ranks to devices with it) */
int eph_b200_device_count(int *out);
int eph_b200_create(const eph_b200_config *cfg, eph_b200_handle **out);
int eph_b200_destroy(eph_b200_handle *h);
const char *eph_b200_last_error(const eph_b200_handle *h);
/* message of a failed create() (no handle exists yet) */
const char *eph_b200_create_error(void);

/* Replaces EPH_Beta's spline set (eph_beta.h:96-125, :164-198) and
 * EPH_Spline::operator() (eph_spline.h:134-142).  Coefficients are the
 * reference's {a,b,c,d} per interval in absolute x, built on the host.
 *   coeff_rho_r_sq [n_elements][n_rho][4],  coeff_alpha / coeff_beta [n_elements][n_beta][4]  */
int eph_b200_set_tables(eph_b200_handle *h, int n_elements, int n_rho, double inv_dr_sq, const double *coeff_rho_r_sq,
                        int n_beta, double inv_drho, const double *coeff_alpha, const double *coeff_beta,
                        double r_cutoff_sq, double rho_cutoff);
/* rho(r) splines in r (eph_beta.h:157-162; knots spaced dr, [n_elements][n_rho][4]) for eph_model 2 (PRB), the one
 * place the reference evaluates them per step (fix_eph.cpp:530).  Call after set_tables. */
int eph_b200_set_rho_r_table(eph_b200_handle *h, int n_elements, int n_rho, double inv_dr, const double *coeff_rho_r);

/* Replaces EPH_FDM's constructors and state vectors (eph_fdm.h:28-119, :415-446).
 * box = {x0,x1,y0,y1,z0,z1}; fields are [nz][ny][nx] (index i + j*nx + k*nx*ny);
 * flag: 0 constant, 1 dynamic, 2 zero-derivative wall; t_dyn: 1 = C_e, kappa_e follow T_e. */
int eph_b200_set_grid(eph_b200_handle *h, int nx, int ny, int nz, const double *box, int steps, const double *T_e,
                      const double *S_e, const double *rho_e, const double *C_e, const double *kappa_e,
                      const int16_t *flag, const uint16_t *t_dyn);
/* Temperature dependent parameters (eph_fdm.h:74-104): C_e(T), kappa_e(T) spline
 * coefficients [n_T][4] and the E_e(T) running-sum table [n_T] (EPH_Linear). */
int eph_b200_set_grid_tables(eph_b200_handle *h, int n_T, double dT, const double *coeff_C_e_T,
                             const double *coeff_kappa_e_T, const double *E_e_T);
/* which: 0 T_e 1 S_e 2 rho_e 3 C_e 4 kappa_e 5 dT_e ; out/in are host arrays of nx*ny*nz doubles */
int eph_b200_get_grid(eph_b200_handle *h, int which, double *out);
int eph_b200_put_grid(eph_b200_handle *h, int which, const double *in);
/* EPH_FDM::get_T_total, eph_fdm.h:189-196 */
int eph_b200_mean_T(eph_b200_handle *h, double *out);
/* sub-steps used by the last solve (eph_fdm.h:302-313) */
int eph_b200_last_substeps(eph_b200_handle *h, int *out);

/* Replaces FixEPH::reset_dt and EPH_FDM::set_dt (fix_eph.cpp:909-916, eph_fdm.h:155-158) */
int eph_b200_set_dt(eph_b200_handle *h, double dt, double boltz);

/* `fix eph/coloured/exp` (fix_eph_coloured_exp.cpp): eph_model 4 with an exponential memory kernel of time constant
 * tau0 (its arg[5]) on both forces, f <- f_prev (1 - zeta) + zeta f_new, zeta = 1 - exp(-dt / tau0) (:190, :563-569,
 * :619-625, :686), applied to group atoms with rho_i > 0; the filtered forces are what is added to f, deposited into the
 * grid and reported per atom.  tau0 <= 0 switches the filter off.  get/set_colour_state move the filter's per-atom state
 * (f_dis, f_sto: [nlocal][3] each), which migrates with the atoms (pack_exchange / unpack_exchange / copy_arrays,
 * :793-825): new storage starts from zero, set_atoms keeps the values of a registration of the same size, and a caller
 * whose atoms were re-ordered registers them again after set_atoms. */
int eph_b200_set_colour(eph_b200_handle *h, double tau0);
int eph_b200_get_colour_state(eph_b200_handle *h, double *f_dis, double *f_sto, int memspace);
int eph_b200_set_colour_state(eph_b200_handle *h, const double *f_dis, const double *f_sto, int memspace);

/* Gather records of the two list sweeps.  packed_records = 1 (default; EPH_B200_RECORDS=packed): steps that walk the
 * inner list read 16-byte fixed-point positions (42-bit fractions of a period P, a power of two > 2 (r_c + 2 inner_skin);
 * quantum P / 2^42 = 3.6e-12 A at the defaults) and 16-byte block-floating-point vectors (v, u, z: 40-bit mantissas,
 * <= 1.8e-12 of the largest component) -- half the 32-byte sectors per list slot of the fp64 records.  The error this
 * adds to rho_i, w_i, f_EPH, f_RNG stays below ~1e-11 of the largest value (the bar is 1e-10).
 * packed_records = 0 (EPH_B200_RECORDS=exact): fp64 records everywhere (agreement with the reference ~1e-14).
 * Packed records need model 4, at most four elements and the inner list; otherwise the fp64 records are used.
 * get_precision reports what the next step will use (period and position quantum in A, 0 when fp64). */
int eph_b200_set_precision(eph_b200_handle *h, int packed_records);
int eph_b200_get_precision(eph_b200_handle *h, int *packed_records, double *period, double *position_quantum);

/* LAMMPS' neighbor->skin and the skin of the device-side inner list (two-level Verlet list: the sweeps walk a
 * list cut at r_c + inner_skin that the device rebuilds from LAMMPS' list; a device-side displacement check falls
 * back to LAMMPS' list whenever the inner one could be incomplete, so results never depend on this setting).
 * inner_skin = 0 disables the inner list, negative keeps the current value (default 0.4 A, or EPH_B200_INNER_SKIN).
 * Call it before set_neighbors: with a list already registered the sweeps walk LAMMPS' list until the next
 * set_neighbors applies the new setting (the age of the registered list is unknown to the engine). */
int eph_b200_set_skin(eph_b200_handle *h, double skin, double inner_skin);
/* how often the inner list was (re)built and how many steps saw it invalidated */
int eph_b200_list_stats(eph_b200_handle *h, long long *inner_builds, long long *fallback_steps);

/* Per-atom data that only changes when LAMMPS re-neighbours (atom->type, mask, tag
 * for nlocal+nghost atoms) and the ghost->owner map of a single-rank periodic
 * run (what Comm::forward_comm(Fix*) would realise through
 * pack/unpack_forward_comm, fix_eph.cpp:951-1009). ghost_owner may be NULL when nghost==0. */
int eph_b200_set_atoms(eph_b200_handle *h, int nlocal, int nghost, const int *type, const int *mask,
                       const int64_t *tag, const int *ghost_owner, int memspace);

/* Replaces FixEPH::init_list + the per-step use of list->numneigh/firstneigh
 * (fix_eph.cpp:289-291, :437-448).  Call only when neighbor->ago == 0.
 * Entries may carry LAMMPS' special-bond bits; they are masked with NEIGHMASK. */
int eph_b200_set_neighbors_csr(eph_b200_handle *h, int nlocal, const int64_t *offsets, const int *neigh, int memspace);
int eph_b200_set_neighbors_lammps(eph_b200_handle *h, int nlocal, const int *numneigh, int *const *firstneigh);

/* Alternative to uploading LAMMPS' list: build the same full list (rows of local atoms over locals and ghosts,
 * pairs closer than cutoff = r_c + neighbor->skin) on the device from the positions x [nlocal+nghost][3].  Call when
 * neighbor->ago == 0, after set_atoms.  x == NULL: the positions the engine keeps itself (after eph_b200_resident_upload).
 * get_neighbors reads the list in use back (offsets [nlocal+1], neigh). */
int eph_b200_build_neighbors(eph_b200_handle *h, const double *x, double cutoff, int memspace);
int eph_b200_get_neighbors(eph_b200_handle *h, int64_t *offsets, int *neigh, long long *n_entries);

/* Replaces FixEPH::post_force (fix_eph.cpp:841-907) for model PRL:
 * xi generation, calculate_environment (:431-466), the three ghost broadcasts,
 * force_prl (:687-837) and f += f_EPH (+ f_RNG).
 * x, v: [nlocal+nghost][3]; f: [nlocal][3] read-modify-write.
 * xi_inject: [nlocal][3] Gaussians to use instead of the built-in counter-based
 * stream (parity testing), or NULL.  ntimestep keys the built-in stream. */
int eph_b200_post_force(eph_b200_handle *h, const double *x, const double *v, double *f, const double *xi_inject,
                        long long ntimestep, int memspace);

/* post_force in two halves for multi-rank runs.  begin: xi, the density pass (rho_i and the pair sums W_i of owned
 * atoms).  Then ONE ghost exchange moves {rho, Wx, Wy, Wz} from owners to ghosts -- it replaces the reference's
 * forward comms RHO and WI (fix_eph.cpp:870-871, :743-744); XI (:863-864) needs no exchange because ghosts regenerate
 * the owner's Gaussians from the atom tag.  end: per-atom coupling, the force pass, f += f_EPH (+ f_RNG).
 * pack/unpack move the payload between the engine and caller-owned DEVICE buffers of n*4 doubles (the transport is
 * NCCL or peer memory, e.g. torch.distributed); ghosts with ghost_owner[g] >= 0 in set_atoms are images of this
 * rank's own atoms and are filled internally, ghosts with ghost_owner[g] < 0 must be covered by unpack. */
int eph_b200_post_force_begin(eph_b200_handle *h, const double *x, const double *v, const double *xi_inject,
                              long long ntimestep, int memspace);
int eph_b200_pack_ghost_payload(eph_b200_handle *h, int n, const int *send_index_dev, double *buf_dev);
int eph_b200_unpack_ghost_payload(eph_b200_handle *h, int n, const int *recv_index_dev, const double *buf_dev);
int eph_b200_post_force_end(eph_b200_handle *h, double *f, int memspace);

/* end_of_step in two halves for multi-rank runs: begin deposits this rank's energy into the grid source term, the
 * caller all-reduces that array over ranks (the reference's MPI_Allreduce, eph_fdm.h:481), end solves the grid on
 * every rank (no broadcast needed).  bind_grid_source makes the engine use a caller-owned DEVICE array of
 * nx*ny*nz doubles as the source term so the transport can reduce it in place. */
int eph_b200_end_of_step_begin(eph_b200_handle *h, const double *x, const double *v, int memspace);
int eph_b200_end_of_step_end(eph_b200_handle *h, double *E_local);
int eph_b200_bind_grid_source(eph_b200_handle *h, double *dT_e_dev);
/* Sharded grid solve (multi-rank alternative to the redundant solve above; replaces MPI_Allreduce + rank-0 solve +
 * MPI_Bcast of eph_fdm.h:481-491 with all-reduce + slab solve with halo planes + all-gather).  Every rank keeps the
 * whole grid in memory and updates only its slab of z-planes [z_begin, z_end); the stencil reads the two neighbouring
 * planes (periodic in z) from the same full-size array, so between two sub-steps the caller copies those planes from
 * the ranks that own them into its own T_e array in place (ncclSend/ncclRecv on grid_device_ptr(h, 0, ...), which
 * follows the double buffer: query it after every sub-step), and after the last sub-step all-gathers the slabs.
 *   end_of_step_begin; all-reduce of the source term; grid_plan_substeps(&n);
 *   n x { grid_substep(z_begin, z_end); halo exchange (not after the last one) }; all-gather of T_e;
 *   end_of_step_end_external
 * plan_substeps refreshes temperature-dependent cells and returns the reference's sub-step count (eph_fdm.h:290-313;
 * 0 without flag 4); the last planned sub-step also clears the source term on all planes.  Work is enqueued on the
 * grid stream when one is set.  grid_device_ptr: which as in get_grid (0 T_e, current buffer; 5 source term). */
int eph_b200_grid_plan_substeps(eph_b200_handle *h, int *n_substeps);
int eph_b200_grid_substep(eph_b200_handle *h, int z_begin, int z_end);
int eph_b200_end_of_step_end_external(eph_b200_handle *h, double *E_local);
int eph_b200_grid_device_ptr(eph_b200_handle *h, int which, double **ptr);
/* ---- Multi-rank data plane inside the engine (NCCL over NVLink / NVSwitch) --------------------------------------------
 * One rank per GPU along LAMMPS' spatial decomposition.  With a communicator attached and a ghost map registered the
 * plain entry points do the whole multi-rank step themselves:
 *   eph_b200_post_force    density pass -> ghost exchange -> coupling -> force pass.  The exchange replaces the
 *                          reference's forward comms RHO and WI (fix_eph.cpp:870-871, :743-744) by ONE grouped
 *                          ncclSend/ncclRecv of {rho, Wx, Wy, Wz} per peer; XI (:863-864) travels only when the caller
 *                          injects Gaussians (the built-in stream is keyed on atom tags: ghosts regenerate it).
 *   eph_b200_end_of_step   deposit -> ncclAllReduce of the source term (MPI_Allreduce, eph_fdm.h:481) -> solve on every
 *                          rank (no MPI_Bcast, eph_fdm.h:490), or with set_grid_sharding the z-slab solve with halo
 *                          planes by ncclSend/ncclRecv between sub-steps and one ncclAllGather.
 *   E_local stays this rank's contribution (the fix sums it over ranks like fix_eph.cpp:402).
 * comm_get_id: on rank 0, 128 bytes (ncclUniqueId) for the caller to broadcast (MPI_Bcast in the fix);
 * comm_init: collective over all ranks, on the handle's device.  NCCL is bound at run time (libnccl.so.2); without it
 * these calls fail with EPH_B200_ERR_COMM -- there is no host fall-back.
 * set_ghost_map (host arrays; call after every set_atoms): for each peer rank the owned atoms it holds as ghosts
 * (send_index: local indices, concatenated in peer order, in the order the peer expects them) and the ghost slots
 * (recv_slot: nlocal + g) that peer fills, in the order it sends them.  Ghosts that are images of the rank's own atoms
 * stay with ghost_owner >= 0 in set_atoms.  exchange_ghosts / reduce_and_solve are the two collective halves for
 * callers that drive post_force_begin/_end and end_of_step_begin themselves.
 * Transport of the two per-step ghost exchanges ({rho, W} here, {x, v} in eph_b200_refresh_ghosts): when every rank of
 * the communicator can map every other rank's memory (one node, NVLink / PCIe peer access) comm_init sets up windows
 * (cudaIpc) and the exchanges become one kernel that stores the rows straight into the receivers' memory plus one that
 * scatters them (csrc/eph_p2p.cuh); otherwise grouped ncclSend / ncclRecv.  The choice is made once, by all ranks
 * together; EPH_B200_EXCHANGE=nccl forces send / receive, EPH_B200_P2P_WINDOW_MB (default 256) sizes the window.  With
 * peer memory set_ghost_map is collective (one small all-reduce): ghost rows that do not fit a rank's share of a
 * window make it fail on every rank together, with the size it needs.
 * comm_transport: 0 no communicator, 1 NCCL send / receive, 2 peer memory.  A peer that fails to arrive within 5 s
 * (EPH_B200_P2P_TIMEOUT_MS) sets bit 8 of the status word instead of hanging the device, and the next end_of_step
 * that returns the energy fails with EPH_B200_ERR_COMM. */
#define EPH_B200_COMM_ID_BYTES 128
int eph_b200_comm_get_id(void *id128);
int eph_b200_comm_init(eph_b200_handle *h, const void *id128, int rank, int nranks);
int eph_b200_comm_transport(const eph_b200_handle *h);
int eph_b200_set_ghost_map(eph_b200_handle *h, int npeers, const int *peer_rank, const int *send_count, const int *send_index,
                           const int *recv_count, const int *recv_slot);
int eph_b200_exchange_ghosts(eph_b200_handle *h);
int eph_b200_set_grid_sharding(eph_b200_handle *h, int on);
int eph_b200_reduce_and_solve(eph_b200_handle *h, double *E_local);

/* Optional second CUDA stream for the grid.  When set, end_of_step_begin makes that stream wait for the deposit,
 * end_of_step_end enqueues the solve on it, and the engine's main stream waits for the solve only where it needs
 * T_e or dT_e again (the force pass, the next deposit, grid read-backs).  A caller that issues its all-reduce of the
 * source term on the same stream thus overlaps all-reduce + solve with the next step's density pass. */
int eph_b200_set_grid_stream(eph_b200_handle *h, void *stream);
/* Optional communication stream and boundary-first density pass (multi-rank overlap of the one ghost exchange a step
 * needs; the reference serialises its forward comms, fix_eph.cpp:743-744, :863-871, with the sweeps).
 * set_boundary_atoms names the owned atoms other ranks hold as ghosts (local indices, after set_atoms): the density
 * pass then sweeps the tiles holding them first.  With a communication stream set, pack_ghost_payload and
 * unpack_ghost_payload run on it -- pack waits only for those boundary tiles, post_force_end waits for unpack -- so a
 * caller that issues its all-to-all on the same stream hides the exchange behind the sweep of the interior tiles. */
int eph_b200_set_comm_stream(eph_b200_handle *h, void *stream);
int eph_b200_set_boundary_atoms(eph_b200_handle *h, int n, const int *index, int memspace);

/* Replaces FixEPH::end_of_step (fix_eph.cpp:350-429): energy bookkeeping,
 * EPH_FDM::insert_energy (eph_fdm.h:172-179), EPH_FDM::solve (:267-400) and the
 * 8-column per-atom output.  E_local (host pointer, may be NULL) receives this
 * rank's energy transfer of the step; passing NULL avoids the host sync.
 * x may be NULL: the positions of the last post_force are used (the Verlet loop
 * does not move atoms between post_force and end_of_step), saving their upload. */
int eph_b200_end_of_step(eph_b200_handle *h, const double *x, const double *v, double *E_local, int memspace);

/* Replaces the integrator hooks (fix_eph.cpp:305-348). mass_by_type is 1-based [ntypes+1]. */
int eph_b200_initial_integrate(eph_b200_handle *h, double *x, double *v, const double *f, const double *mass_by_type,
                               double dtv, double dtf, int memspace);
int eph_b200_final_integrate(eph_b200_handle *h, double *v, const double *f, const double *mass_by_type, double dtf,
                             int memspace);

/* Ghost atoms following their owners on the device: what LAMMPS' Comm::forward_comm() does for x and v (the fix sets
 * comm->ghost_velocity, fix_eph.cpp:82) when those arrays live on the GPU.  Images of the rank's own atoms are shifted
 * copies (set_atoms' ghost_owner >= 0); ghosts owned by other ranks arrive over the NCCL ghost map.  The first call after
 * set_atoms, while the ghost coordinates are still LAMMPS' own, records the image shifts; later calls apply them.
 * x, v: DEVICE arrays [nlocal + nghost][3]. */
int eph_b200_refresh_ghosts(eph_b200_handle *h, double *x, double *v);

/* Device-resident integration (replaces the host round trips around FixEPH::initial_integrate /
 * post_force / final_integrate / end_of_step, fix_eph.cpp:305-429, when LAMMPS' own arrays are host arrays).  The engine
 * keeps x, v of all atoms and f of the local ones between the hooks.  Per step only the pair forces go up; x (after the
 * drift: the pair style and the neighbour check need it), f (after post_force) and v (after the second kick, optional)
 * come down.  Valid while this fix is the integrator of all atoms it is given and the last fix to change f.
 *   resident_upload             after every set_atoms (re-neighbouring): x, v [nlocal + nghost][3] from the host once
 *   resident_initial_integrate  kick + drift on the device with the forces of the last resident_post_force (f: host
 *                               forces for the very first step, else may be NULL); ghosts follow; x_out <- x[nlocal][3].
 *                               start_post_force_step >= 0 (the time step; only when LAMMPS will not re-neighbour in this
 *                               step and the Gaussians are the built-in stream): the first half of post_force -- records
 *                               and density pass, which need x and v only -- is started at once and runs while x travels
 *                               and the host computes its pair forces; resident_post_force then continues from there
 *   resident_post_force         f: host forces of the other contributors in; f_out <- the same plus f_EPH (+ f_RNG)
 *                               (may be f itself; NULL: not needed on the host this step)
 *   resident_final_integrate    second kick; v_out <- v[nlocal][3] (NULL: not needed on the host this step)
 *   resident_end_of_step        end_of_step on the resident velocities */
int eph_b200_resident_upload(eph_b200_handle *h, const double *x, const double *v);
int eph_b200_resident_initial_integrate(eph_b200_handle *h, const double *f, const double *mass_by_type, double dtv, double dtf,
                                        double *x_out, long long start_post_force_step);
int eph_b200_resident_post_force(eph_b200_handle *h, const double *f, double *f_out, const double *xi_inject, long long ntimestep);
int eph_b200_resident_final_integrate(eph_b200_handle *h, const double *mass_by_type, double dtf, double *v_out);
int eph_b200_resident_end_of_step(eph_b200_handle *h, double *E_local);
/* which: 0 x, 1 v, 2 f of the local atoms -> HOST out [nlocal][3] (e.g. v after the first kick on a step in which LAMMPS
 * re-neighbours: it migrates and re-orders the atoms from its host arrays) */
int eph_b200_resident_get(eph_b200_handle *h, int which, double *out);

/* FixEPH::array (fix_eph.cpp:406-428): [nlocal][8] = rho, beta(rho), f_EPH xyz, f_RNG xyz */
int eph_b200_get_peratom(eph_b200_handle *h, double *array8, int memspace);
/* probes for parity tests; LAMMPS order.
 * which: 0 rho_i[nlocal+nghost] 1 w_i[nlocal][3] 2 xi_i[nlocal][3] 3 f_EPH[nlocal][3] 4 f_RNG[nlocal][3] */
int eph_b200_get_probe(eph_b200_handle *h, int which, double *out);

/* Replaces FixEPH::pack_forward_comm / unpack_forward_comm (fix_eph.cpp:951-1009)
 * for a host transport (LAMMPS' MPI comm): gathers / scatters the payload of
 * `state` between device-resident per-atom arrays and a host buffer, between
 * post_force_begin and post_force_end.  RHO: rho (1 double per atom); WI: the
 * pair sums W of the density pass (3 doubles; the receiver forms w itself);
 * XI: injected Gaussians (3 doubles).  pack returns the number of doubles written. */
int eph_b200_pack_forward(eph_b200_handle *h, int state, int n, const int *list, double *buf);
int eph_b200_unpack_forward(eph_b200_handle *h, int state, int n, int first, const double *buf);

/* Per-kernel device timing with CUDA events on the launch stream (benchmarks).  kernel_times returns the number of
 * distinct kernels seen since profiling was switched on and fills up to `max` entries (total ms, launches).
 * on: 0 off, 1 every kernel (two events per launch: about 3 us of stream time each), 2 only the two list sweeps. */
int eph_b200_set_profiling(eph_b200_handle *h, int on);
int eph_b200_kernel_times(eph_b200_handle *h, int max, const char **names, double *ms, long long *counts);

/* blocks until everything enqueued so far has finished */
int eph_b200_synchronize(eph_b200_handle *h);
/* number of kernels this handle has launched since creation */
long long eph_b200_launch_count(const eph_b200_handle *h);
/* device status word: bit0 rho_i > rho_cutoff seen (eph_beta.h:174-180), bit1 T_e clamped at 0 (eph_fdm.h:391-394),
 * bit2 non-finite force, bit8 a peer-memory ghost exchange timed out */
int eph_b200_status_word(eph_b200_handle *h, unsigned *out);

#ifdef __cplusplus
}
#endif
#endif /* EPH_B200_H */
