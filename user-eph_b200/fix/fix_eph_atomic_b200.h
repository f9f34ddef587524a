/* fix eph/atomic/b200 -- B200-native drop-in for `fix eph/atomic` (LLNL/USER-EPH, fix_eph_atomic.cpp).
 *
 * Same command line as the reference fix (fix_eph_atomic.cpp:39-56):
 *   fix ID group eph/atomic/b200 seed flags T_e T_infile inner_loops T_out beta_file kappa_file elem...
 * optionally followed by keyword pairs the reference does not have:
 *   rng mars|philox   source of the Gaussians xi_i (default philox: counter-based on the device, keyed on atom tags;
 *                     mars: LAMMPS' RanMars on the host in the reference's order, :808-816, uploaded every step)
 *   device N          CUDA device ordinal (default 0)
 *   comm device|lammps ghost values through the engine's own owner map (one rank per box, default there) or through LAMMPS'
 *                     Comm::forward_comm(Fix*) with host buffers in the reference's order (default for several ranks)
 * The same hooks are registered (:301-311), the same outputs produced (f_ID[1] = electronic energy of the group,
 * f_ID[2] = its mean temperature, 12 per-atom columns) and the per-atom electronic energy migrates with its atom
 * (copy_arrays / pack_exchange / unpack_exchange, :939-955).  All per-timestep work is done by libeph_b200
 * (include/eph_b200_atomic.h).
 * Build with -DEPH_B200_REPLACE_FIX_EPH to register under the name `eph/atomic` itself.
 */
#ifdef FIX_CLASS
#ifdef EPH_B200_REPLACE_FIX_EPH
FixStyle(eph/atomic,FixEPHAtomicB200)
#else
FixStyle(eph/atomic/b200,FixEPHAtomicB200)
#endif
#else

#ifndef LMP_FIX_EPH_ATOMIC_B200_H
#define LMP_FIX_EPH_ATOMIC_B200_H

#include <cstdint>
#include <string>
#include <vector>

#include "fix.h"

#include "eph_b200_atomic.h"
#include "eph_kappa_tables.h"
#include "eph_tables.h"

namespace LAMMPS_NS {

class FixEPHAtomicB200 : public Fix {
 public:
  enum class FixState : unsigned int { NONE, RHO, XI, WI, EI, OWNER };   // fix_eph_atomic.h:35-41 + the owner map
  // FixEPHAtomic::Flag (fix_eph_atomic.h:44-51)
  enum Flag : int { FRICTION = 0x01, RANDOM = 0x02, HEAT = 0x04, NOINT = 0x08, NOFRICTION = 0x10, NORANDOM = 0x20 };

  FixEPHAtomicB200(class LAMMPS *, int, char **);
  ~FixEPHAtomicB200() override;

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
  void post_run() override {}
  int pack_forward_comm(int, int *, double *, int, int *) override;
  void unpack_forward_comm(int, int, double *) override;
  int pack_exchange(int, double *) override;
  int unpack_exchange(int, double *) override;
  void copy_arrays(int, int, int) override;

  // test-driver views (tests/lammps_shim/fix_driver.h)
  void probe_copy(int which, size_t nlocal, size_t ntotal, double *out);
  size_t grid_size() const { return 0; }
  void grid_T(double *) {}
  void set_energy_host(const double *E);   // overwrite E_a_i of the local atoms (tests: start from a gradient)

 protected:
  int myID, nrPS;
  FixState state;
  int eph_flag;
  int types;
  std::vector<int> type_map_beta, type_map_kappa;
  eph_b200::BetaTables beta;
  eph_b200::KappaTables kappa;
  eph_b200_atomic_handle *dev;

  double dtv, dtf;
  double r_cutoff;
  int inner_loops;
  int seed;
  class RanMars *random;
  bool rng_mars;
  bool comm_lammps;
  class NeighList *list;
  double Ee, Te;
  size_t n;

  double **array;   // [nmax][12] per-atom output (array_atom)
  double *E_a_i;    // [nmax] per-atom electronic energy, migrates with the atoms; the device copy is refreshed from it
                    // whenever LAMMPS may have re-ordered the atoms (re-neighbouring) and written back every step
  std::vector<double> xi_host;
  std::vector<int64_t> tag64;   // atom->tag widened to 64 bits for the C ABI (tagint may be 32 bits wide)
  std::vector<int> ghost_owner;
  std::vector<int64_t> csr_offsets;
  std::vector<int> csr_neigh;
  long long atoms_epoch;
  bool need_upload;

  void upload_topology();
  void forward(FixState st);
  void check(int rc, const char *what);
};

}  // namespace LAMMPS_NS
#endif
#endif
