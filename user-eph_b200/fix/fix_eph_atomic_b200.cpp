// fix eph/atomic/b200: host side of the B200-native `fix eph/atomic`.  Keeps the reference's command syntax, file
// formats, hook order, outputs and error messages (reference fix_eph_atomic.cpp) and forwards the per-timestep work to
// libeph_b200 (include/eph_b200_atomic.h).
#include "fix_eph_atomic_b200.h"

#include <mpi.h>

#include "eph_device_select.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neigh_list.h"
#include "neigh_request.h"
#include "neighbor.h"
#include "random_mars.h"
#include "update.h"

using namespace LAMMPS_NS;
using namespace FixConst;

/* arguments: identical positions to FixEPHAtomic (fix_eph_atomic.cpp:39-56)
 *  3 seed | 4 flags | 5 T_e | 6 T_infile | 7 inner loops | 8 T_out | 9 beta file | 10 kappa file | 11.. element per type
 *  then optional keyword pairs */
FixEPHAtomicB200::FixEPHAtomicB200(LAMMPS *lmp, int narg, char **arg)
    : Fix(lmp, narg, arg), dev(nullptr), random(nullptr), array(nullptr), E_a_i(nullptr) {
  if (narg < 12) error->all(FLERR, "fix_eph_atomic: too few arguments");
  if (atom->natoms < 1) error->all(FLERR, "fix_eph_atomic: error no atoms in simulation");
  MPI_Comm_rank(world, &myID);
  MPI_Comm_size(world, &nrPS);

  state = FixState::NONE;
  vector_flag = 1;
  size_vector = 2;
  global_freq = 1;
  extvector = 1;
  nevery = 1;
  peratom_flag = 1;
  size_peratom_cols = 12;
  peratom_freq = 1;
  comm_forward = 3;
  comm->ghost_velocity = 1;

  seed = atoi(arg[3]);
  random = new RanMars(lmp, seed + myID);

  eph_flag = strtol(arg[4], NULL, 0);
  if (myID == 0) {
    std::cout << '\n' << "Flag read: " << arg[4] << " -> " << eph_flag << '\n';
    if (eph_flag & Flag::FRICTION) std::cout << "Friction evaluation: ON\n";
    if (eph_flag & Flag::RANDOM) std::cout << "Random evaluation: ON\n";
    if (eph_flag & Flag::HEAT) std::cout << "Heat diffusion solving: ON\n";
    if (eph_flag & Flag::NOINT) std::cout << "No integration: ON\n";
    if (eph_flag & Flag::NOFRICTION) std::cout << "No friction application: ON\n";
    if (eph_flag & Flag::NORANDOM) std::cout << "No random application: ON\n";
    std::cout << '\n';
  }

  const double v_Te = atof(arg[5]);
  // arg[6] (initial temperatures from a file) is accepted and ignored, as in the reference (:212-224)
  inner_loops = atoi(arg[7]);
  if (inner_loops < 1) inner_loops = 0;

  const int n_elem = 11;
  types = atom->ntypes;
  if (types > (narg - n_elem)) error->all(FLERR, "fix_eph_atomic: number of types larger than provided in fix");

  try {
    beta = eph_b200::load_beta_file(arg[9]);
    kappa = eph_b200::load_kappa_file(arg[10]);
  } catch (const std::exception &e) {
    error->all(FLERR, e.what());
  }
  if (beta.n_elements < 1) error->all(FLERR, "fix_eph_atomic: no elements found in beta file");
  if (kappa.n_elements < 1) error->all(FLERR, "fix_eph_atomic: no elements found in kappa file");
  r_cutoff = beta.r_cutoff;

  type_map_beta.assign(types, -1);
  type_map_kappa.assign(types, -1);
  for (int i = 0; i < types; ++i) {
    type_map_beta[i] = beta.find(arg[n_elem + i]);
    type_map_kappa[i] = kappa.find(arg[n_elem + i]);
    if (type_map_beta[i] < 0 || type_map_kappa[i] < 0) error->all(FLERR, "fix_eph_atomic: elements not found in input file");
  }

  rng_mars = false;
  comm_lammps = nrPS > 1;
  int device = -1;
  for (int k = n_elem + types; k < narg; ++k) {
    const bool is_keyword = strcmp(arg[k], "rng") == 0 || strcmp(arg[k], "device") == 0 || strcmp(arg[k], "comm") == 0;
    if (!is_keyword) continue;   // an extra element name (the reference's decks list more names than types)
    if (k + 1 >= narg) error->all(FLERR, "fix eph/atomic/b200: keyword without a value");
    const char *val = arg[k + 1];
    if (strcmp(arg[k], "rng") == 0) {
      if (strcmp(val, "mars") == 0) rng_mars = true;
      else if (strcmp(val, "philox") == 0) rng_mars = false;
      else error->all(FLERR, "fix eph/atomic/b200: rng must be mars or philox");
    } else if (strcmp(arg[k], "comm") == 0) {
      if (strcmp(val, "lammps") == 0) comm_lammps = true;
      else if (strcmp(val, "device") == 0) comm_lammps = false;
      else error->all(FLERR, "fix eph/atomic/b200: comm must be device or lammps");
    } else {
      device = atoi(val);
    }
    ++k;
  }
  if (nrPS > 1 && !comm_lammps)
    error->all(FLERR, "fix eph/atomic/b200: comm device needs one rank per box (ghosts must be images of the rank's own atoms); use comm lammps");

  dtv = update->dt;
  dtf = 0.5 * update->dt * force->ftm2v;

  list = nullptr;
  n = 0;
  atoms_epoch = -1;
  need_upload = true;
  grow_arrays(atom->nmax);
  atom->add_callback(0);
  const size_t ntotal = (size_t)atom->nlocal + atom->nghost;
  std::fill_n(&(array[0][0]), size_peratom_cols * ntotal, 0.0);
  std::fill_n(E_a_i, ntotal, 0.0);

  // per-atom energies from the initial temperature, their sum and the mean temperature (:212-253)
  Ee = 0.0;
  Te = 0.0;
  int atom_counter = 0;
  for (int i = 0; i < atom->nlocal; ++i)
    if (atom->mask[i] & groupbit) {
      const int ek = type_map_kappa[atom->type[i] - 1];
      E_a_i[i] = eph_b200::linear_eval(kappa.E_T[ek], v_Te);
      Ee += E_a_i[i];
      const double T = eph_b200::linear_reverse(kappa.E_T[ek], E_a_i[i]);
      Te += T;
      atom_counter++;
      array[i][9] = E_a_i[i];    // populate_array before the first step: every other column is still zero
      array[i][11] = T;
      array[i][1] = eph_b200::CubicTable(beta.beta[type_map_beta[atom->type[i] - 1]])(0.0);   // beta(rho_i = 0), :408
    }
  // over all ranks like the reference: the energies add up, the temperature is the mean over the ranks that hold atoms of
  // the group of their local means (fix_eph_atomic.cpp:236-260)
  if (atom_counter > 0) Te /= static_cast<double>(atom_counter);
  int proc_counter = atom_counter > 0 ? 1 : 0;
  MPI_Allreduce(MPI_IN_PLACE, &Ee, 1, MPI_DOUBLE, MPI_SUM, world);
  MPI_Allreduce(MPI_IN_PLACE, &Te, 1, MPI_DOUBLE, MPI_SUM, world);
  MPI_Allreduce(MPI_IN_PLACE, &proc_counter, 1, MPI_INT, MPI_SUM, world);
  Te /= static_cast<double>(proc_counter);

  // the device engine
  eph_b200_atomic_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.device = device >= 0 ? device : eph_b200::default_device(myID);   // node-local rank modulo the visible devices
  cfg.ntypes = types;
  cfg.type_map_beta = type_map_beta.data();
  cfg.type_map_kappa = type_map_kappa.data();
  cfg.groupbit = groupbit;
  cfg.flags = eph_flag;
  cfg.seed = (unsigned long long)seed;
  cfg.inner_loops = inner_loops;
  cfg.stream = nullptr;
  if (eph_b200_atomic_create(&cfg, &dev) != EPH_B200_OK) error->all(FLERR, eph_b200_atomic_create_error());
  {
    std::vector<double> t_rho = eph_b200::BetaTables::flatten(beta.rho_r_sq);
    std::vector<double> t_alpha = eph_b200::BetaTables::flatten(beta.alpha);
    std::vector<double> t_beta = eph_b200::BetaTables::flatten(beta.beta);
    check(eph_b200_atomic_set_beta_tables(dev, beta.n_elements, (int)beta.n_rho, beta.inv_dr_sq(), t_rho.data(), (int)beta.n_beta,
                                          beta.inv_drho(), t_alpha.data(), t_beta.data(), beta.r_cutoff_sq, beta.rho_cutoff),
          "set_beta_tables");
    std::vector<double> k_rho = eph_b200::BetaTables::flatten(kappa.rho_r_sq);
    std::vector<double> k_E = eph_b200::KappaTables::flatten(kappa.E_T);
    std::vector<double> k_K = eph_b200::KappaTables::flatten(kappa.K_T);
    check(eph_b200_atomic_set_kappa_tables(dev, kappa.n_elements, kappa.n_pairs, (int)kappa.n_r, kappa.rho_r_sq.at(0).inv_dx,
                                           k_rho.data(), kappa.r_cutoff_sq, (int)kappa.n_T, kappa.dT, k_E.data(), k_K.data()),
          "set_kappa_tables");
  }
  check(eph_b200_atomic_set_dt(dev, update->dt, force->boltz), "set_dt");
  check(eph_b200_atomic_set_comm_mode(dev, comm_lammps ? 1 : 0), "set_comm_mode");
}

FixEPHAtomicB200::~FixEPHAtomicB200() {
  delete random;
  atom->delete_callback(id, 0);
  memory->destroy(array);
  memory->destroy(E_a_i);
  eph_b200_atomic_destroy(dev);
}

void FixEPHAtomicB200::check(int rc, const char *what) {
  if (rc != EPH_B200_OK) {
    std::string msg = std::string("fix eph/atomic/b200: ") + what + ": " + eph_b200_atomic_last_error(dev);
    error->all(FLERR, msg);
  }
}

void FixEPHAtomicB200::init() {
  if (domain->dimension == 2) error->all(FLERR, "Cannot use fix eph with 2d simulation");
  if (domain->nonperiodic != 0) error->all(FLERR, "Cannot use nonperiodic boundares with fix eph");
  if (domain->triclinic) error->all(FLERR, "Cannot use fix eph with triclinic box");
  // full neighbour list including ghosts, cut-off r_c of the beta file (fix_eph_atomic.cpp:286-288)
  int request_style = NeighConst::REQ_FULL | NeighConst::REQ_GHOST;
  auto req = neighbor->add_request(this, request_style);
  req->set_cutoff(r_cutoff);
  reset_dt();
}

void FixEPHAtomicB200::init_list(int, NeighList *ptr) {
  this->list = ptr;
  need_upload = true;
}

int FixEPHAtomicB200::setmask() {
  int mask = 0;
  mask |= POST_FORCE;
  mask |= END_OF_STEP;
  mask |= INITIAL_INTEGRATE;
  mask |= FINAL_INTEGRATE;
  return mask;
}

// velocity-Verlet half steps on LAMMPS' host arrays (fix_eph_atomic.cpp:313-359)
void FixEPHAtomicB200::initial_integrate(int) {
  if (eph_flag & Flag::NOINT) return;
  double **x = atom->x, **v = atom->v, **f = atom->f;
  const double *mass = atom->mass;
  const int *type = atom->type, *mask = atom->mask;
  const int nlocal = atom->nlocal;
  for (int i = 0; i < nlocal; ++i) {
    if (!(mask[i] & groupbit)) continue;
    const double dtfm = dtf / mass[type[i]];
    for (int d = 0; d < 3; ++d) v[i][d] += dtfm * f[i][d];
    for (int d = 0; d < 3; ++d) x[i][d] += dtv * v[i][d];
  }
}

void FixEPHAtomicB200::final_integrate() {
  if (eph_flag & Flag::NOINT) return;
  double **v = atom->v, **f = atom->f;
  const double *mass = atom->mass;
  const int *type = atom->type, *mask = atom->mask;
  const int nlocal = atom->nlocal;
  for (int i = 0; i < nlocal; ++i) {
    if (!(mask[i] & groupbit)) continue;
    const double dtfm = dtf / mass[type[i]];
    for (int d = 0; d < 3; ++d) v[i][d] += dtfm * f[i][d];
  }
}

// What only changes when LAMMPS re-neighbours: types / masks / tags, the ghost->owner map, the list, and -- because
// LAMMPS may have sorted or migrated atoms since the last step -- the per-atom energies in their new order.
void FixEPHAtomicB200::upload_topology() {
  const int nlocal = atom->nlocal, nghost = atom->nghost;
  ghost_owner.assign(nghost, -1);
  if (!comm_lammps) {
    forward(FixState::OWNER);   // ghost -> owner: one forward comm of the owner's local index through our own pack/unpack
    for (int g = 0; g < nghost; ++g)
      if (ghost_owner[g] < 0 || ghost_owner[g] >= nlocal) error->all(FLERR, "fix eph/atomic/b200: ghost atom without a local owner");
  }
  // LAMMPS' tagint is 32 or 64 bits wide depending on the build (-DLAMMPS_SMALLBIG, the default, has 32): widen here
  tag64.resize((size_t)nlocal + nghost);
  for (size_t i = 0; i < tag64.size(); ++i) tag64[i] = static_cast<int64_t>(atom->tag[i]);
  check(eph_b200_atomic_set_atoms(dev, nlocal, nghost, atom->type, atom->mask, tag64.data(),
                                  comm_lammps ? nullptr : ghost_owner.data(), EPH_B200_HOST),
        "set_atoms");
  if (!list) error->all(FLERR, "fix eph/atomic/b200: no neighbour list");
  csr_offsets.assign((size_t)nlocal + 1, 0);
  for (int i = 0; i < nlocal; ++i) csr_offsets[i + 1] = csr_offsets[i] + list->numneigh[i];
  csr_neigh.resize((size_t)csr_offsets[nlocal]);
  for (int i = 0; i < nlocal; ++i) std::copy(list->firstneigh[i], list->firstneigh[i] + list->numneigh[i], csr_neigh.begin() + csr_offsets[i]);
  check(eph_b200_atomic_set_neighbors_csr(dev, nlocal, csr_offsets.data(), csr_neigh.data(), EPH_B200_HOST), "set_neighbors");
  if (nlocal > 0) check(eph_b200_atomic_set_energy(dev, E_a_i, EPH_B200_HOST), "set_energy");
  atoms_epoch = ((long long)nlocal << 32) | (unsigned)nghost;
  need_upload = false;
}

void FixEPHAtomicB200::post_force(int) {
  const int nlocal = atom->nlocal, nghost = atom->nghost;
  const long long epoch = ((long long)nlocal << 32) | (unsigned)nghost;
  if (need_upload || neighbor->ago == 0 || epoch != atoms_epoch) upload_topology();

  const double *xi = nullptr;
  if ((eph_flag & Flag::RANDOM) && rng_mars) {   // the reference's stream (fix_eph_atomic.cpp:808-816)
    xi_host.assign(3 * (size_t)nlocal, 0.0);
    const int *mask = atom->mask;
    for (int i = 0; i < nlocal; ++i)
      if (mask[i] & groupbit) {
        xi_host[3 * (size_t)i + 0] = random->gaussian();
        xi_host[3 * (size_t)i + 1] = random->gaussian();
        xi_host[3 * (size_t)i + 2] = random->gaussian();
      }
    xi = xi_host.data();
  }
  if (!comm_lammps) {
    if (nlocal + nghost == 0) return;
    check(eph_b200_atomic_post_force(dev, &atom->x[0][0], &atom->v[0][0], nlocal ? &atom->f[0][0] : nullptr, xi, update->ntimestep,
                                     EPH_B200_HOST),
          "post_force");
    return;
  }
  // the reference's transport and order of forward comms (fix_eph_atomic.cpp:803-825, :549-550); every rank takes part
  // in every comm, with or without atoms
  double dummy[3] = {0, 0, 0};
  const double *xp = (nlocal + nghost) ? &atom->x[0][0] : dummy, *vp = (nlocal + nghost) ? &atom->v[0][0] : dummy;
  check(eph_b200_atomic_post_force_begin(dev, xp, vp, xi, update->ntimestep, EPH_B200_HOST), "post_force");
  forward(FixState::EI);
  if (eph_flag & Flag::RANDOM) forward(FixState::XI);
  forward(FixState::RHO);
  check(eph_b200_atomic_post_force_mid(dev), "post_force");
  if (eph_flag & Flag::FRICTION) forward(FixState::WI);
  check(eph_b200_atomic_post_force_end(dev, nlocal ? &atom->f[0][0] : nullptr, EPH_B200_HOST), "post_force");
}

void FixEPHAtomicB200::forward(FixState st) {
  state = st;
  comm->forward_comm(this);
  state = FixState::NONE;
}

void FixEPHAtomicB200::end_of_step() {
  const int nlocal = atom->nlocal;
  double E = 0.0, T = 0.0;
  if (!comm_lammps) {
    check(eph_b200_atomic_end_of_step(dev, &E, &T), "end_of_step");
  } else {
    const int loops = eph_b200_atomic_heat_loops(dev);   // heat_solve with the EI forward comm of every loop (:720-721)
    for (int it = 0; it < loops; ++it) {
      check(eph_b200_atomic_heat_begin(dev), "heat_begin");
      forward(FixState::EI);
      check(eph_b200_atomic_heat_end(dev), "heat_end");
    }
    check(eph_b200_atomic_summary(dev, &E, &T), "summary");
  }
  // the reference's reductions (fix_eph_atomic.cpp:382-397): energies add up, temperatures are averaged over the ranks
  // that hold group atoms (T comes back as NaN from a rank without any)
  int proc_counter = (T == T) ? 1 : 0;
  if (!proc_counter) T = 0.0;
  MPI_Allreduce(MPI_IN_PLACE, &E, 1, MPI_DOUBLE, MPI_SUM, world);
  MPI_Allreduce(MPI_IN_PLACE, &T, 1, MPI_DOUBLE, MPI_SUM, world);
  MPI_Allreduce(MPI_IN_PLACE, &proc_counter, 1, MPI_INT, MPI_SUM, world);
  Ee = E;
  Te = T / static_cast<double>(proc_counter);
  if (nlocal > 0) {
    check(eph_b200_atomic_get_energy(dev, E_a_i, EPH_B200_HOST), "get_energy");   // E_a_i travels with the atoms on the host
    check(eph_b200_atomic_get_peratom(dev, &array[0][0], EPH_B200_HOST), "get_peratom");
  }
}

void FixEPHAtomicB200::reset_dt() {
  dtv = update->dt;
  dtf = 0.5 * update->dt * force->ftm2v;
  check(eph_b200_atomic_set_dt(dev, update->dt, force->boltz), "set_dt");
}

void FixEPHAtomicB200::grow_arrays(int ngrow) {
  n = ngrow;
  memory->grow(array, ngrow, size_peratom_cols, "eph:array");
  memory->grow(E_a_i, ngrow, "eph:E_a_i");
  array_atom = array;
}

double FixEPHAtomicB200::compute_vector(int i) {   // fix_eph_atomic.cpp:839-847
  if (i == 1) return Te;
  return Ee;
}

// FixEPHAtomic::pack_forward_comm / unpack_forward_comm (fix_eph_atomic.cpp:849-927): the payloads live on the device
static int abi_state(FixEPHAtomicB200::FixState st) {
  using S = FixEPHAtomicB200::FixState;
  return st == S::RHO ? 1 : st == S::XI ? 2 : st == S::WI ? 3 : st == S::EI ? 4 : 0;
}

int FixEPHAtomicB200::pack_forward_comm(int n, int *list, double *data, int, int *) {
  int m = 0;
  if (state == FixState::OWNER) {
    const int nlocal = atom->nlocal;
    for (int i = 0; i < n; ++i) {   // LAMMPS forwards ghosts of ghosts in later swaps: resolve through the part already known
      const int src = list[i];
      data[m++] = static_cast<double>(src < nlocal ? src : ghost_owner[src - nlocal]);
    }
  } else if (abi_state(state)) {
    m = eph_b200_atomic_pack_forward(dev, abi_state(state), n, list, data);
    if (m < 0) check(m, "pack_forward");
  }
  return m;
}

void FixEPHAtomicB200::unpack_forward_comm(int n, int first, double *data) {
  if (state == FixState::OWNER) {
    const int nlocal = atom->nlocal;
    for (int i = 0; i < n; ++i) ghost_owner[first + i - nlocal] = static_cast<int>(data[i]);
  } else if (abi_state(state)) {
    check(eph_b200_atomic_unpack_forward(dev, abi_state(state), n, first, data), "unpack_forward");
  }
}

// the per-atom electronic energy migrates with its atom (fix_eph_atomic.cpp:939-955)
int FixEPHAtomicB200::pack_exchange(int i, double *buf) {
  buf[0] = E_a_i[i];
  return 1;
}

int FixEPHAtomicB200::unpack_exchange(int nlocal, double *buf) {
  E_a_i[nlocal] = buf[0];
  need_upload = true;
  return 1;
}

void FixEPHAtomicB200::copy_arrays(int i, int j, int) {
  E_a_i[j] = E_a_i[i];
  need_upload = true;
}

double FixEPHAtomicB200::memory_usage() { return (double)n * (size_peratom_cols + 1) * sizeof(double); }

void FixEPHAtomicB200::set_energy_host(const double *E) {
  std::copy(E, E + atom->nlocal, E_a_i);
  need_upload = true;
}

void FixEPHAtomicB200::probe_copy(int which, size_t, size_t, double *out) {
  if (atoms_epoch < 0) {   // nothing registered on the device yet: the constructor's state
    if (which == 6) { std::copy(E_a_i, E_a_i + atom->nlocal + atom->nghost, out); return; }
    error->all(FLERR, "fix eph/atomic/b200: probe before the first step");
  }
  check(eph_b200_atomic_get_probe(dev, which, out), "get_probe");
}
