/* fix eph/b200 -- B200-native drop-in for `fix eph` (LLNL/USER-EPH).
 *
 * Same command line as the reference fix (fix_eph.cpp:36-58):
 *   fix ID group eph/b200 seed flags model rho_e C_e kappa_e T_e NX NY NZ T_infile freq T_out beta_file elem...
 * optionally followed by keyword pairs the reference does not have:
 *   rng mars|philox   source of the Gaussians xi_i (default philox: counter-based, generated on the device and keyed
 *                     on atom tags, so no XI ghost exchange is needed; mars: LAMMPS' RanMars on the host in the
 *                     reference's order, fix_eph.cpp:854-861, uploaded every step)
 *   device N          CUDA device ordinal (default: rank modulo visible devices)
 *   neigh lammps|device the full neighbour list comes from LAMMPS (uploaded when it is rebuilt; default) or is built
 *                     on the device from the positions (LAMMPS then builds no list for this fix)
 *   comm device|lammps|nccl  ghost values through the engine's own owner map (single rank, default), through LAMMPS'
 *                     Comm::forward_comm(Fix*) with host buffers and MPI_Allreduce of the grid source term, as the
 *                     reference does (default for several ranks), or -- one rank per GPU -- through the engine's own
 *                     NCCL data plane (one grouped send/recv of {rho, W} per peer, ncclAllReduce of the source term)
 *   integrate host|device  the velocity-Verlet half steps of the fix run on LAMMPS' host arrays (default) or on the
 *                     device, where x, v, f then stay between the hooks: per step only the pair forces go up, and x
 *                     (after the drift), f (after post_force) and v (after the second kick) come down.  Needs the fix
 *                     group to be all atoms, this fix to be the last one that changes f, and a re-neighbouring schedule
 *                     known in advance (neigh_modify ... check no): on those steps v is brought down after the first
 *                     kick as well, because LAMMPS migrates and re-orders the atoms from its host arrays
 *   sync N            with integrate device: atom->f and atom->v on the host are brought up to date every N-th step only
 *                     (default 1; set it to the thermo / dump interval), x every step
 *   grid replicated|sharded  with comm nccl: every rank solves the whole grid (default) or only its z-slab, with halo
 *                     planes between sub-steps and one all-gather (replaces the reference's MPI_Bcast, eph_fdm.h:490)
 * The same hooks are registered (fix_eph.cpp:293-302) and the same outputs are produced
 * (f_ID[1], f_ID[2], 8 per-atom columns); all per-timestep work is done by libeph_b200 (include/eph_b200.h).
 * Build with -DEPH_B200_REPLACE_FIX_EPH to register under the name `eph` itself.
 *
 * The same class serves `fix eph/coloured/exp` (fix_eph_coloured_exp.cpp), the reference's fork of model 4 with an
 * exponential memory kernel on both forces: under a style name containing "coloured" arg[5] is the time constant tau0
 * instead of the model number (fix_eph_coloured_exp.cpp:43), the filter runs on the device (eph_b200_set_colour) and
 * its per-atom state (f_dis, f_sto) migrates with the atoms through pack_exchange / unpack_exchange / copy_arrays.
 */
#ifdef FIX_CLASS
#ifdef EPH_B200_REPLACE_FIX_EPH
FixStyle(eph,FixEPHB200)
FixStyle(eph/coloured/exp,FixEPHB200)
#else
FixStyle(eph/b200,FixEPHB200)
FixStyle(eph/coloured/exp/b200,FixEPHB200)
#endif
#else

#ifndef LMP_FIX_EPH_B200_H
#define LMP_FIX_EPH_B200_H

#include <cstdint>
#include <string>
#include <vector>

#include "fix.h"

#include "eph_b200.h"
#include "eph_grid_io.h"
#include "eph_tables.h"

namespace LAMMPS_NS {

class FixEPHB200 : public Fix {
 public:
  // same enumerations as FixEPH (fix_eph.h:35-60)
  enum class FixState : unsigned int { NONE, RHO, XI, WI, OWNER };
  enum Flag : int { FRICTION = 0x01, RANDOM = 0x02, FDM = 0x04, NOINT = 0x08, NOFRICTION = 0x10, NORANDOM = 0x20 };
  enum Model : int { TESTING = -1, NONE = 0, TTM = 1, PRB = 2, PRLCM = 3, PRL = 4 };

  FixEPHB200(class LAMMPS *, int, char **);
  ~FixEPHB200() override;

  void init() override;
  void init_list(int id, class NeighList *ptr) override;
  int setmask() override;
  void initial_integrate(int) override;
  void post_force(int) override;
  void final_integrate() override;
  void end_of_step() override;
  void reset_dt() override;
  void grow_arrays(int) override;
  double compute_vector(int) override;
  double memory_usage() override;
  void post_run() override;
  int pack_forward_comm(int, int *, double *, int, int *) override;
  void unpack_forward_comm(int, int, double *) override;
  int pack_exchange(int, double *) override;
  int unpack_exchange(int, double *) override;
  void copy_arrays(int, int, int) override;

  // read-only views used by the test driver (tests/lammps_shim/fix_driver.h)
  void probe_copy(int which, size_t nlocal, size_t ntotal, double *out);
  size_t grid_size() const { return grid.ncell(); }
  void grid_T(double *out);

 protected:
  static constexpr size_t max_file_length = 256;

  int myID, nrPS;
  FixState state;
  int eph_flag, eph_model;
  int types;
  std::vector<int> type_map;

  eph_b200::BetaTables beta;   // host copy of the tables (parsed from arg[16])
  eph_b200::GridState grid;    // host description of the FDM grid (parameters, file names)
  eph_b200_handle *dev;        // the device engine

  double dtv, dtf;
  double r_cutoff, r_cutoff_sq, rho_cutoff;
  int T_freq;
  char T_out[max_file_length];
  char T_state[max_file_length];
  double eta_factor;
  int seed;
  class RanMars *random;
  bool rng_mars;
  bool comm_lammps;
  bool comm_nccl;               // keyword `comm nccl`: the engine exchanges ghosts / sums the grid source over NCCL itself
  bool grid_sharded;            // keyword `grid sharded`: with comm nccl every rank advances only its z-slab of the grid
  bool neigh_device;
  bool integrate_device;        // keyword `integrate device`: x, v, f of the atoms stay on the device between the hooks
  int sync_every;               // keyword `sync N`: with integrate device, LAMMPS' host f and v are refreshed every N-th step
  double extra_skin;            // keyword `extra_skin X`: the list is requested X A longer than r_c + neighbor->skin (see init())
  int peratom_every;            // keyword `peratom N`: array_atom is refreshed every N-th step (0: never)
  class NeighList *list;
  double Ee;
  size_t n;

  bool coloured;                // style eph/coloured/exp: exponential memory kernel on both forces
  double tau0;
  double **f_sto_i, **f_dis_i;  // coloured: filtered random / friction force of the last step, [nmax][3], migrate with the atoms
  double **array;               // [nmax][8] per-atom output (array_atom)
  std::vector<double> xi_host;  // rng mars: Gaussians of this step
  std::vector<int64_t> tag64;   // atom->tag widened to 64 bits for the C ABI (tagint may be 32 bits wide)
  std::vector<int> ghost_owner; // owner's local index of each ghost ...
  std::vector<int> ghost_rank;  // ... and the rank that owns it
  std::vector<double> source_buf; // comm lammps on several ranks: the grid source term on its way through MPI_Allreduce
  std::vector<double> owner_buf;
  long long atoms_epoch;        // (nlocal,nghost) signature of the last upload
  bool need_upload;
  long long v_synced_step;      // integrate device: the step whose half-kicked v went down to the host (a re-neighbouring step)

  void upload_topology();
  void check(int rc, const char *what);
};

}  // namespace LAMMPS_NS
#endif
#endif
