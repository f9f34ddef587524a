// fix eph/b200: host side of the B200-native `fix eph`.  Keeps the reference's
// command syntax, file formats, hook order, outputs and error messages
// (reference fix_eph.cpp) and forwards the per-timestep work to libeph_b200.
#include "fix_eph_b200.h"

#include <mpi.h>

#include "eph_device_select.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
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

/* arguments: identical positions to FixEPH (fix_eph.cpp:36-58)
 *  3 seed | 4 flags | 5 model | 6 rho_e | 7 C_e | 8 kappa_e | 9 T_e | 10-12 NX NY NZ | 13 T_infile | 14 freq
 *  15 T_out | 16 beta file | 17.. element per type | then optional keyword pairs */
FixEPHB200::FixEPHB200(LAMMPS *lmp, int narg, char **arg)
    : Fix(lmp, narg, arg), dev(nullptr), random(nullptr), coloured(false), tau0(0.0), f_sto_i(nullptr), f_dis_i(nullptr) {
  if (narg < 18) error->all(FLERR, "Illegal fix eph command: too few arguments");
  coloured = strstr(arg[2], "coloured") != nullptr;   // fix eph/coloured/exp: arg[5] is tau0 (fix_eph_coloured_exp.cpp:43)
  if (atom->natoms < 1) error->all(FLERR, "fix_eph: error no atoms in simulation");
  MPI_Comm_rank(world, &myID);
  MPI_Comm_size(world, &nrPS);

  state = FixState::NONE;

  vector_flag = 1;
  size_vector = 2;
  global_freq = 1;
  extvector = 1;
  nevery = 1;
  peratom_flag = 1;
  size_peratom_cols = 8;
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
    if (eph_flag & Flag::FDM) std::cout << "FDM grid solving: ON\n";
    if (eph_flag & Flag::NOINT) std::cout << "No integration: ON\n";
    if (eph_flag & Flag::NOFRICTION) std::cout << "No friction application: ON\n";
    if (eph_flag & Flag::NORANDOM) std::cout << "No random application: ON\n";
    std::cout << '\n';
  }
  time_integrate = (eph_flag & Flag::NOINT) ? 0 : 1;

  if (coloured) {
    tau0 = atof(arg[5]);
    if (!(tau0 > 0.0)) error->all(FLERR, "fix eph/coloured/exp/b200: tau0 must be positive");
    eph_model = Model::PRL;
    maxexchange = 6;
    if (myID == 0) std::cout << "\nColoured noise, exponential kernel: tau0 = " << tau0 << " (B200 device path)\n" << std::endl;
  } else {
    eph_model = atoi(arg[5]);
    if (myID == 0) std::cout << "\nModel read: " << arg[5] << " -> " << eph_model << " (B200 device path)\n" << std::endl;
  }
  if (eph_model == Model::PRLCM)
    error->all(FLERR, "fix eph/b200: model 3 (PRLCM) is not offered: the reference indexes its rho(r) table with jtype - i "
                      "there (fix_eph.cpp:601)");
  if (eph_model != Model::PRL && eph_model != Model::NONE && eph_model != Model::TTM && eph_model != Model::PRB)
    error->all(FLERR, "fix eph/b200: unknown model (1 TTM, 2 PRB, 4 PRL run on the device)");

  const double v_rho = atof(arg[6]);
  const double v_Ce = atof(arg[7]);
  const double v_kappa = atof(arg[8]);
  const double v_Te = atof(arg[9]);
  const int nx = atoi(arg[10]);
  const int ny = atoi(arg[11]);
  const int nz = atoi(arg[12]);

  try {
    if (strcmp("NULL", arg[13]) == 0) {
      if (nx < 1 || ny < 1 || nz < 1) error->all(FLERR, "FixEPH: non-positive grid values");
      grid = eph_b200::make_uniform_grid(nx, ny, nz, domain->boxlo, domain->boxhi, v_Te, v_Ce, v_rho, v_kappa);
      strcpy(T_state, "T.restart");
    } else {
      grid = eph_b200::load_grid_file(arg[13]);
      snprintf(T_state, max_file_length, "%s.restart", arg[13]);
    }
  } catch (const std::exception &e) {
    error->all(FLERR, e.what());
  }

  T_freq = atoi(arg[14]);
  T_out[0] = '\0';
  if (T_freq > 0) snprintf(T_out, max_file_length, "%s", arg[15]);

  types = atom->ntypes;
  if (types > (narg - 17)) error->all(FLERR, "Fix eph: number of types larger than provided in fix");

  try {
    beta = eph_b200::load_beta_file(arg[16]);
  } catch (const std::exception &e) {
    error->all(FLERR, e.what());
  }
  if (beta.n_elements < 1) error->all(FLERR, "Fix eph: no elements found in input file");
  r_cutoff = beta.r_cutoff;
  r_cutoff_sq = beta.r_cutoff_sq;
  rho_cutoff = beta.rho_cutoff;

  type_map.assign(types, -1);
  for (int i = 0; i < types; ++i) {
    type_map[i] = beta.find(arg[17 + i]);
    if (type_map[i] < 0) error->all(FLERR, "Fix eph: elements not found in input file");
  }

  // optional keyword pairs after the element names
  rng_mars = false;
  comm_lammps = nrPS > 1;
  comm_nccl = false;
  grid_sharded = false;
  neigh_device = false;
  integrate_device = false;
  sync_every = 1;
  extra_skin = 0.0;
  peratom_every = 1;
  int device = -1;
  // token by token: decks list more element names than atom types (e.g. `Ni.beta Ni Ni` with one type); like the
  // reference those extras are skipped, and a keyword is recognised wherever it stands
  for (int k = 17 + types; k < narg; ++k) {
    const bool is_keyword = strcmp(arg[k], "rng") == 0 || strcmp(arg[k], "device") == 0 || strcmp(arg[k], "neigh") == 0 ||
                            strcmp(arg[k], "peratom") == 0 || strcmp(arg[k], "comm") == 0 || strcmp(arg[k], "grid") == 0 ||
                            strcmp(arg[k], "integrate") == 0 || strcmp(arg[k], "sync") == 0 || strcmp(arg[k], "extra_skin") == 0;
    if (!is_keyword) continue;   // an extra element name
    if (k + 1 >= narg) error->all(FLERR, "fix eph/b200: keyword without a value");
    const char *val = arg[k + 1];
    if (strcmp(arg[k], "rng") == 0) {
      if (strcmp(val, "mars") == 0) rng_mars = true;
      else if (strcmp(val, "philox") == 0) rng_mars = false;
      else error->all(FLERR, "fix eph/b200: rng must be mars or philox");
    } else if (strcmp(arg[k], "device") == 0) {
      device = atoi(val);
    } else if (strcmp(arg[k], "neigh") == 0) {
      if (strcmp(val, "device") == 0) neigh_device = true;
      else if (strcmp(val, "lammps") == 0) neigh_device = false;
      else error->all(FLERR, "fix eph/b200: neigh must be device or lammps");
    } else if (strcmp(arg[k], "peratom") == 0) {
      // the 8 per-atom columns live on the device; bringing them to array_atom costs 64 bytes per atom over PCIe, so
      // the cadence is the user's: every N-th step (default 1 = the reference's behaviour, 0 = never)
      peratom_every = atoi(val);
      if (peratom_every < 0) error->all(FLERR, "fix eph/b200: peratom must be >= 0");
    } else if (strcmp(arg[k], "integrate") == 0) {
      if (strcmp(val, "device") == 0) integrate_device = true;
      else if (strcmp(val, "host") == 0) integrate_device = false;
      else error->all(FLERR, "fix eph/b200: integrate must be host or device");
    } else if (strcmp(arg[k], "sync") == 0) {
      sync_every = atoi(val);
      if (sync_every < 1) error->all(FLERR, "fix eph/b200: sync must be >= 1");
    } else if (strcmp(arg[k], "extra_skin") == 0) {
      extra_skin = atof(val);
      if (!(extra_skin >= 0.0)) error->all(FLERR, "fix eph/b200: extra_skin must be >= 0");
    } else if (strcmp(arg[k], "grid") == 0) {
      if (strcmp(val, "sharded") == 0) grid_sharded = true;
      else if (strcmp(val, "replicated") == 0) grid_sharded = false;
      else error->all(FLERR, "fix eph/b200: grid must be replicated or sharded");
    } else {   // comm
      if (strcmp(val, "lammps") == 0) { comm_lammps = true; comm_nccl = false; }
      else if (strcmp(val, "device") == 0) { comm_lammps = false; comm_nccl = false; }
      else if (strcmp(val, "nccl") == 0) { comm_lammps = false; comm_nccl = true; }
      else error->all(FLERR, "fix eph/b200: comm must be device, lammps or nccl");
    }
    ++k;
  }

  if (integrate_device && (eph_flag & Flag::NOINT)) error->all(FLERR, "fix eph/b200: integrate device contradicts flag 8 (no integration)");
  if (integrate_device && comm_lammps) error->all(FLERR, "fix eph/b200: integrate device needs comm device (one rank) or comm nccl");

  eta_factor = sqrt(2.0 * force->boltz / update->dt);
  dtv = update->dt;
  dtf = 0.5 * update->dt * force->ftm2v;

  // the device engine
  eph_b200_config cfg;
  memset(&cfg, 0, sizeof cfg);
  if (device < 0) device = eph_b200::default_device(myID);   // node-local rank modulo the visible devices
  cfg.device = device;
  cfg.ntypes = types;
  cfg.type_map = type_map.data();
  cfg.groupbit = groupbit;
  cfg.flags = eph_flag;
  cfg.model = eph_model;
  cfg.seed = (unsigned long long)seed;
  cfg.rank = myID;
  cfg.nranks = nrPS;
  cfg.stream = nullptr;
  if (eph_b200_create(&cfg, &dev) != EPH_B200_OK) error->all(FLERR, eph_b200_create_error());
  if (comm_nccl && nrPS > 1) {
    // the engine's own transport (NCCL over NVLink): rank 0 creates the communicator id, MPI carries its 128 bytes
    char id[EPH_B200_COMM_ID_BYTES];
    memset(id, 0, sizeof id);
    int ok = 1;
    if (myID == 0) ok = eph_b200_comm_get_id(id) == EPH_B200_OK ? 1 : 0;
    MPI_Bcast(&ok, 1, MPI_INT, 0, world);
    if (!ok) error->all(FLERR, std::string("fix eph/b200: comm nccl: ") + eph_b200_create_error());
    MPI_Bcast(id, EPH_B200_COMM_ID_BYTES, MPI_CHAR, 0, world);
    check(eph_b200_comm_init(dev, id, myID, nrPS), "comm_init");
    check(eph_b200_set_grid_sharding(dev, grid_sharded ? 1 : 0), "set_grid_sharding");
  }

  {
    std::vector<double> t_rho = eph_b200::BetaTables::flatten(beta.rho_r_sq);
    std::vector<double> t_alpha = eph_b200::BetaTables::flatten(beta.alpha);
    std::vector<double> t_beta = eph_b200::BetaTables::flatten(beta.beta);
    check(eph_b200_set_tables(dev, beta.n_elements, (int)beta.n_rho, beta.inv_dr_sq(), t_rho.data(), (int)beta.n_beta,
                              beta.inv_drho(), t_alpha.data(), t_beta.data(), beta.r_cutoff_sq, beta.rho_cutoff),
          "set_tables");
    if (eph_model == Model::PRB) {   // the one model that evaluates rho(r) per step (fix_eph.cpp:530)
      std::vector<double> t_rho_r = eph_b200::BetaTables::flatten(beta.rho_r);
      check(eph_b200_set_rho_r_table(dev, beta.n_elements, (int)beta.n_rho, beta.rho_r.at(0).inv_dx, t_rho_r.data()),
            "set_rho_r_table");
    }
  }
  check(eph_b200_set_grid(dev, (int)grid.nx, (int)grid.ny, (int)grid.nz, grid.box, (int)grid.steps, grid.T_e.data(),
                          grid.S_e.data(), grid.rho_e.data(), grid.C_e.data(), grid.kappa_e.data(), grid.flag.data(),
                          grid.t_dyn.data()),
        "set_grid");
  if (grid.has_tables)
    check(eph_b200_set_grid_tables(dev, (int)grid.C_e_T.size(), grid.E_e_T.dx, grid.C_e_T.k.data(), grid.kappa_e_T.k.data(),
                                   grid.E_e_T.y.data()),
          "set_grid_tables");
  check(eph_b200_set_dt(dev, update->dt, force->boltz), "set_dt");
  check(eph_b200_set_skin(dev, neighbor->skin + extra_skin, -1.0), "set_skin");
  if (coloured) check(eph_b200_set_colour(dev, tau0), "set_colour");

  array = nullptr;
  list = nullptr;
  n = 0;
  atoms_epoch = -1;
  need_upload = true;
  v_synced_step = -1;

  grow_arrays(atom->nmax);
  atom->add_callback(0);
  std::fill_n(&(array[0][0]), size_peratom_cols * (size_t)(atom->nlocal + atom->nghost), 0);
  if (coloured) {   // fix_eph_coloured_exp.cpp:229-230
    std::fill_n(&(f_sto_i[0][0]), 3 * (size_t)(atom->nlocal + atom->nghost), 0.);
    std::fill_n(&(f_dis_i[0][0]), 3 * (size_t)(atom->nlocal + atom->nghost), 0.);
  }

  Ee = 0.0;
}

FixEPHB200::~FixEPHB200() {
  delete random;
  atom->delete_callback(id, 0);
  memory->destroy(array);
  memory->destroy(f_sto_i);
  memory->destroy(f_dis_i);
  eph_b200_destroy(dev);
}

void FixEPHB200::check(int rc, const char *what) {
  if (rc != EPH_B200_OK) {
    std::string msg = std::string("fix eph/b200: ") + what + ": " + eph_b200_last_error(dev);
    error->all(FLERR, msg);
  }
}

void FixEPHB200::init() {
  if (domain->dimension == 2) error->all(FLERR, "Cannot use fix eph with 2d simulation");
  if (domain->nonperiodic != 0) error->all(FLERR, "Cannot use nonperiodic boundares with fix eph");
  if (domain->triclinic) error->all(FLERR, "Cannot use fix eph with triclinic box");

  // full neighbour list including ghosts, cut-off r_c (fix_eph.cpp:273-275); with `neigh device` the engine builds
  // the same list from the positions and LAMMPS need not build (nor the host upload) a second full list for this fix
  if (!neigh_device) {
    int request_style = NeighConst::REQ_FULL | NeighConst::REQ_GHOST;
    auto req = neighbor->add_request(this, request_style);
    // `extra_skin X`: the engine's inner list (skin 0.4 A) can be rebuilt from LAMMPS' aged list only while twice the
    // displacement since LAMMPS' build plus the inner skin still fits into the list's skin; LAMMPS re-neighbours at
    // half ITS skin, so for the last fifth of a list's life the sweeps would have to walk the full list.  A list
    // requested X = 0.4 A longer closes that gap; the pair set the forces see is unchanged (r_c decides).
    req->set_cutoff(r_cutoff + extra_skin);
  }

  // a `neighbor` / `neigh_modify` command may have changed the skin since the fix line or between runs; the engine's
  // inner list is validated against it on the device, so it must know the current value before every run
  check(eph_b200_set_skin(dev, neighbor->skin + extra_skin, -1.0), "set_skin");
  need_upload = true;
  if (integrate_device && neighbor->dist_check)
    error->all(FLERR, "fix eph/b200: integrate device needs a re-neighbouring schedule known in advance (neigh_modify every N delay 0 check no)");

  reset_dt();
}

void FixEPHB200::init_list(int, NeighList *ptr) {
  this->list = ptr;
  need_upload = true;
}

int FixEPHB200::setmask() {
  int mask = 0;
  mask |= POST_FORCE;
  mask |= END_OF_STEP;
  mask |= INITIAL_INTEGRATE;
  mask |= FINAL_INTEGRATE;
  return mask;
}

// Velocity-Verlet half steps (fix_eph.cpp:305-348).  x, v, f live in LAMMPS'
// host arrays here, so these streaming loops stay on the host; the device
// variants (eph_b200_initial_integrate / final_integrate) serve GPU-resident atoms.
void FixEPHB200::initial_integrate(int) {
  if (eph_flag & Flag::NOINT) return;
  if (integrate_device && !need_upload) {
    const int nlocal = atom->nlocal;
    if (nlocal == 0) return;
    // Neighbor::decide(), which runs next: ago + 1 reaches a multiple of `every` that is not below `delay`
    const int ago = neighbor->ago + 1;
    const bool reneigh = neighbor->every > 0 && ago >= neighbor->delay && ago % neighbor->every == 0;
    // without a re-neighbouring ahead (and with the built-in Gaussians) the density pass of this step's post_force starts
    // now and runs while the host computes its pair forces
    const bool early = !reneigh && !((eph_flag & Flag::RANDOM) && rng_mars);
    check(eph_b200_resident_initial_integrate(dev, &atom->f[0][0], atom->mass, dtv, dtf, &atom->x[0][0], early ? (long long)update->ntimestep : -1),
          "initial_integrate");
    // LAMMPS re-neighbours from its host arrays (migration, sorting): on those steps it needs the half-kicked v as well
    if (reneigh) check(eph_b200_resident_get(dev, 1, &atom->v[0][0]), "resident_get");
    v_synced_step = reneigh ? update->ntimestep : -1;
    return;
  }
  v_synced_step = update->ntimestep;   // host loop: LAMMPS' arrays are the authoritative ones this step
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

void FixEPHB200::final_integrate() {
  if (eph_flag & Flag::NOINT) return;
  if (integrate_device) {
    if (atom->nlocal > 0)
      check(eph_b200_resident_final_integrate(dev, atom->mass, dtf, update->ntimestep % sync_every == 0 ? &atom->v[0][0] : nullptr), "final_integrate");
    return;
  }
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

// Everything that only changes when LAMMPS re-neighbours: atom types / masks /
// tags, the ghost->owner map and the neighbour list go to the device once per
// rebuild, never per step.
void FixEPHB200::upload_topology() {
  const int nlocal = atom->nlocal, nghost = atom->nghost;
  if (nrPS > 1 && !comm_lammps && !comm_nccl)
    error->all(FLERR, "fix eph/b200: comm device needs a single rank; use comm nccl (one rank per GPU) or comm lammps");
  ghost_owner.assign(nghost, -1);
  ghost_rank.assign(nghost, -1);
  std::vector<int> peers, send_count, send_index, recv_count, recv_slot;
  if (!comm_lammps) {
    // ghost -> (owner rank, owner's local index): one forward comm through our own pack/unpack
    state = FixState::OWNER;
    comm->forward_comm(this);
    state = FixState::NONE;
    for (int g = 0; g < nghost; ++g)
      if (ghost_rank[g] < 0 || ghost_rank[g] >= nrPS || ghost_owner[g] < 0 || (ghost_rank[g] == myID && ghost_owner[g] >= nlocal))
        error->one(FLERR, "fix eph/b200: ghost atom without an owner");
  }
  if (comm_nccl && nrPS > 1) {
    // The ghost map of the engine's NCCL exchange.  Ghosts owned elsewhere are grouped by owner; every owner is told
    // which of its atoms this rank holds, in this rank's ghost order (the receive order), and answers in kind.
    std::vector<std::vector<int>> want(nrPS), slot(nrPS);
    for (int g = 0; g < nghost; ++g) {
      const int r = ghost_rank[g];
      if (r == myID) continue;             // an image of one of this rank's own atoms: filled inside the engine
      want[r].push_back(ghost_owner[g]);
      slot[r].push_back(nlocal + g);
      ghost_owner[g] = -1;
    }
    std::vector<int> n_want(nrPS), n_asked(nrPS), d_want(nrPS), d_asked(nrPS), flat_want;
    for (int r = 0; r < nrPS; ++r) n_want[r] = (int)want[r].size();
    MPI_Alltoall(n_want.data(), 1, MPI_INT, n_asked.data(), 1, MPI_INT, world);
    int tw = 0, ta = 0;
    for (int r = 0; r < nrPS; ++r) {
      d_want[r] = tw; d_asked[r] = ta;
      tw += n_want[r]; ta += n_asked[r];
      flat_want.insert(flat_want.end(), want[r].begin(), want[r].end());
    }
    std::vector<int> asked(ta > 0 ? ta : 1);
    if (flat_want.empty()) flat_want.push_back(0);
    MPI_Alltoallv(flat_want.data(), n_want.data(), d_want.data(), MPI_INT, asked.data(), n_asked.data(), d_asked.data(), MPI_INT, world);
    for (int r = 0; r < nrPS; ++r) {
      if (r == myID || (n_want[r] == 0 && n_asked[r] == 0)) continue;
      peers.push_back(r);
      send_count.push_back(n_asked[r]);
      send_index.insert(send_index.end(), asked.begin() + d_asked[r], asked.begin() + d_asked[r] + n_asked[r]);
      recv_count.push_back(n_want[r]);
      recv_slot.insert(recv_slot.end(), slot[r].begin(), slot[r].end());
    }
  }
  // LAMMPS' tagint is 32 or 64 bits wide depending on the build (-DLAMMPS_SMALLBIG, the default, has 32): widen here
  tag64.resize((size_t)nlocal + nghost);
  for (size_t i = 0; i < tag64.size(); ++i) tag64[i] = static_cast<int64_t>(atom->tag[i]);
  check(eph_b200_set_atoms(dev, nlocal, nghost, atom->type, atom->mask, tag64.data(),
                           ghost_owner.data(), EPH_B200_HOST),
        "set_atoms");
  if (comm_nccl && nrPS > 1)
    check(eph_b200_set_ghost_map(dev, (int)peers.size(), peers.data(), send_count.data(), send_index.data(), recv_count.data(),
                                 recv_slot.data()),
          "set_ghost_map");
  bool uploaded = false;
  if (integrate_device && neigh_device && nlocal + nghost > 0) {   // x goes up once: the list is built from the resident copy
    for (int i = 0; i < nlocal; ++i)
      if (!(atom->mask[i] & groupbit)) error->one(FLERR, "fix eph/b200: integrate device needs every atom in the fix group");
    check(eph_b200_resident_upload(dev, &atom->x[0][0], &atom->v[0][0]), "resident_upload");
    uploaded = true;
  }
  if (neigh_device) check(eph_b200_build_neighbors(dev, uploaded ? nullptr : &atom->x[0][0], r_cutoff + neighbor->skin + extra_skin, EPH_B200_HOST), "build_neighbors");
  else check(eph_b200_set_neighbors_lammps(dev, nlocal, list->numneigh, list->firstneigh), "set_neighbors");
  // the memory kernel's state in the atoms' present order (LAMMPS may have sorted or migrated them)
  if (coloured && nlocal > 0) check(eph_b200_set_colour_state(dev, &f_dis_i[0][0], &f_sto_i[0][0], EPH_B200_HOST), "set_colour_state");
  if (integrate_device) {
    if (atoms_epoch >= 0 && v_synced_step != update->ntimestep)
      error->all(FLERR, "fix eph/b200: integrate device: LAMMPS re-neighboured on a step the fix did not expect (neigh_modify every N delay 0 check no)");
    for (int i = 0; i < nlocal; ++i)
      if (!(atom->mask[i] & groupbit)) error->one(FLERR, "fix eph/b200: integrate device needs every atom in the fix group");
    if (nlocal + nghost > 0 && !uploaded) check(eph_b200_resident_upload(dev, &atom->x[0][0], &atom->v[0][0]), "resident_upload");
  }
  atoms_epoch = ((long long)nlocal << 32) | (unsigned)nghost;
  need_upload = false;
}

void FixEPHB200::post_force(int) {
  const int nlocal = atom->nlocal, nghost = atom->nghost;
  const long long epoch = ((long long)nlocal << 32) | (unsigned)nghost;
  if (need_upload || neighbor->ago == 0 || epoch != atoms_epoch) upload_topology();

  const double *xi = nullptr;
  if ((eph_flag & Flag::RANDOM) && rng_mars) {
    // the reference's stream: three Gaussians per group atom in local order (fix_eph.cpp:854-861)
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
  if (nlocal + nghost == 0) return;   // no atoms, no ghosts: nobody exchanges anything with this rank
  if (integrate_device) {
    // LAMMPS' own f and v are brought up to date every `sync`-th step (thermo / dump steps); x every step
    double *fp = nlocal ? &atom->f[0][0] : nullptr;
    check(eph_b200_resident_post_force(dev, fp, update->ntimestep % sync_every == 0 ? fp : nullptr, xi, update->ntimestep), "post_force");
    return;
  }
  if (!comm_lammps) {   // one rank, or the engine's NCCL exchange between the two halves of post_force
    check(eph_b200_post_force(dev, &atom->x[0][0], &atom->v[0][0], nlocal ? &atom->f[0][0] : nullptr, xi, update->ntimestep,
                              EPH_B200_HOST),
          "post_force");
    if (coloured && nlocal > 0) check(eph_b200_get_colour_state(dev, &f_dis_i[0][0], &f_sto_i[0][0], EPH_B200_HOST), "get_colour_state");
    return;
  }
  // the reference's transport: ghosts get xi, rho and the w sums through Comm::forward_comm(Fix*) (fix_eph.cpp:863-871, :743-744)
  check(eph_b200_post_force_begin(dev, &atom->x[0][0], &atom->v[0][0], xi, update->ntimestep, EPH_B200_HOST), "post_force");
  if (xi) { state = FixState::XI; comm->forward_comm(this); }
  state = FixState::RHO; comm->forward_comm(this);
  if (eph_flag & Flag::FRICTION) { state = FixState::WI; comm->forward_comm(this); }
  state = FixState::NONE;
  check(eph_b200_post_force_end(dev, nlocal ? &atom->f[0][0] : nullptr, EPH_B200_HOST), "post_force");
  if (coloured && nlocal > 0) check(eph_b200_get_colour_state(dev, &f_dis_i[0][0], &f_sto_i[0][0], EPH_B200_HOST), "get_colour_state");
}

void FixEPHB200::end_of_step() {
  const int nlocal = atom->nlocal;
  double E_local = 0.0;
  const double *xp = nlocal > 0 ? &atom->x[0][0] : nullptr, *vp = nlocal > 0 ? &atom->v[0][0] : nullptr;
  if (integrate_device) {
    if (nlocal > 0 || nrPS > 1) check(eph_b200_resident_end_of_step(dev, &E_local), "end_of_step");
  } else if (nrPS == 1) {
    if (nlocal > 0) check(eph_b200_end_of_step(dev, xp, vp, &E_local, EPH_B200_HOST), "end_of_step");
  } else if (comm_nccl) {
    // every rank, with or without atoms: the engine sums the source term over ranks (ncclAllReduce) and solves
    check(eph_b200_end_of_step(dev, xp, vp, &E_local, EPH_B200_HOST), "end_of_step");
  } else {
    // LAMMPS' transport: every rank deposits its own atoms, the source term is summed over ranks like the reference's
    // sync_before does (eph_fdm.h:479-482), and every rank then solves the same grid (no broadcast needed, :484-491)
    check(eph_b200_end_of_step_begin(dev, xp, vp, EPH_B200_HOST), "end_of_step");
    source_buf.resize(grid.ncell());
    check(eph_b200_get_grid(dev, 5, source_buf.data()), "get_grid");
    MPI_Allreduce(MPI_IN_PLACE, source_buf.data(), (int)source_buf.size(), MPI_DOUBLE, MPI_SUM, world);
    check(eph_b200_put_grid(dev, 5, source_buf.data()), "put_grid");
    check(eph_b200_end_of_step_end(dev, &E_local), "end_of_step");
  }

  // heat map (fix_eph.cpp:397-399); rank 0 only, so failures there must not wait for the other ranks
  if (myID == 0 && T_freq > 0 && (update->ntimestep % T_freq) == 0) {
    std::vector<double> T(grid.ncell());
    if (eph_b200_get_grid(dev, 0, T.data()) != EPH_B200_OK) error->one(FLERR, std::string("fix eph/b200: get_grid: ") + eph_b200_last_error(dev));
    try {
      eph_b200::write_heat_map(grid, T, T_out, (int)(update->ntimestep / T_freq));
    } catch (const std::exception &e) {
      error->one(FLERR, e.what());
    }
  }

  MPI_Allreduce(MPI_IN_PLACE, &E_local, 1, MPI_DOUBLE, MPI_SUM, world);
  Ee += E_local;

  if (nlocal > 0 && peratom_every > 0 && update->ntimestep % peratom_every == 0)
    check(eph_b200_get_peratom(dev, &array[0][0], EPH_B200_HOST), "get_peratom");
}

void FixEPHB200::reset_dt() {
  eta_factor = sqrt(2.0 * force->boltz / update->dt);
  dtv = update->dt;
  dtf = 0.5 * update->dt * force->ftm2v;
  check(eph_b200_set_dt(dev, update->dt, force->boltz), "set_dt");   // also refreshes the filter's zeta factor
}

void FixEPHB200::grow_arrays(int ngrow) {
  n = ngrow;
  memory->grow(array, ngrow, size_peratom_cols, "eph:array");
  array_atom = array;
  if (coloured) {
    memory->grow(f_sto_i, ngrow, 3, "eph:f_sto_i");
    memory->grow(f_dis_i, ngrow, 3, "eph:f_dis_i");
  }
}

// fix eph/coloured/exp: the filtered forces of the last step migrate with their atom (fix_eph_coloured_exp.cpp:793-825)
int FixEPHB200::pack_exchange(int i, double *buf) {
  if (!coloured) return 0;
  int m = 0;
  for (int d = 0; d < 3; ++d) buf[m++] = f_sto_i[i][d];
  for (int d = 0; d < 3; ++d) buf[m++] = f_dis_i[i][d];
  return m;
}

int FixEPHB200::unpack_exchange(int nlocal, double *buf) {
  if (!coloured) return 0;
  int m = 0;
  for (int d = 0; d < 3; ++d) f_sto_i[nlocal][d] = buf[m++];
  for (int d = 0; d < 3; ++d) f_dis_i[nlocal][d] = buf[m++];
  need_upload = true;
  return m;
}

void FixEPHB200::copy_arrays(int i, int j, int) {
  if (!coloured) return;
  for (int d = 0; d < 3; ++d) { f_sto_i[j][d] = f_sto_i[i][d]; f_dis_i[j][d] = f_dis_i[i][d]; }
  need_upload = true;
}

double FixEPHB200::compute_vector(int i) {
  if (i == 0) return Ee;
  if (i == 1) {
    double T = 0.0;
    check(eph_b200_mean_T(dev, &T), "mean_T");
    return T;
  }
  return Ee;
}

int FixEPHB200::pack_forward_comm(int n, int *list, double *data, int, int *) {
  int m = 0;
  if (state == FixState::OWNER) {
    const int nlocal = atom->nlocal;
    // LAMMPS forwards ghosts of ghosts in later swaps: resolve through the part already known
    for (int i = 0; i < n; ++i) {
      const int src = list[i];
      data[m++] = static_cast<double>(src < nlocal ? myID : ghost_rank[src - nlocal]);
      data[m++] = static_cast<double>(src < nlocal ? src : ghost_owner[src - nlocal]);
    }
    return m;
  }
  const int st = state == FixState::RHO ? EPH_B200_STATE_RHO : state == FixState::WI ? EPH_B200_STATE_WI
               : state == FixState::XI ? EPH_B200_STATE_XI : EPH_B200_STATE_NONE;
  if (st != EPH_B200_STATE_NONE) {
    m = eph_b200_pack_forward(dev, st, n, list, data);
    if (m < 0) check(m, "pack_forward");
  }
  return m;
}

void FixEPHB200::unpack_forward_comm(int n, int first, double *data) {
  if (state == FixState::OWNER) {
    const int nlocal = atom->nlocal;
    for (int i = 0; i < n; ++i) {
      ghost_rank[first + i - nlocal] = static_cast<int>(data[2 * i]);
      ghost_owner[first + i - nlocal] = static_cast<int>(data[2 * i + 1]);
    }
    return;
  }
  const int st = state == FixState::RHO ? EPH_B200_STATE_RHO : state == FixState::WI ? EPH_B200_STATE_WI
               : state == FixState::XI ? EPH_B200_STATE_XI : EPH_B200_STATE_NONE;
  if (st != EPH_B200_STATE_NONE) check(eph_b200_unpack_forward(dev, st, n, first, data), "unpack_forward");
}

double FixEPHB200::memory_usage() {
  return (double)n * (size_peratom_cols + (coloured ? 6 : 0)) * sizeof(double);
}

// final grid state, readable as T_infile of a later run (fix_eph.cpp:1019-1021)
void FixEPHB200::post_run() {
  if (myID != 0) return;   // rank 0 only: errors in here are error->one, error->all would wait for ranks that never come
  std::vector<double> T(grid.ncell());
  // temperature-dependent cells carry their last C_e / kappa_e in the reference's restart
  if (eph_b200_get_grid(dev, 0, T.data()) != EPH_B200_OK || eph_b200_get_grid(dev, 3, grid.C_e.data()) != EPH_B200_OK ||
      eph_b200_get_grid(dev, 4, grid.kappa_e.data()) != EPH_B200_OK)
    error->one(FLERR, std::string("fix eph/b200: get_grid: ") + eph_b200_last_error(dev));
  try {
    eph_b200::write_restart(grid, T, T_state);
  } catch (const std::exception &e) {
    error->one(FLERR, e.what());
  }
}

void FixEPHB200::probe_copy(int which, size_t nlocal, size_t, double *out) {
  if (coloured && (which == 5 || which == 6)) {   // the filter state as the host holds it: 5 f_dis, 6 f_sto
    double **src = which == 5 ? f_dis_i : f_sto_i;
    std::copy(&src[0][0], &src[0][0] + 3 * nlocal, out);
    return;
  }
  check(eph_b200_get_probe(dev, which, out), "get_probe");
}

void FixEPHB200::grid_T(double *out) { check(eph_b200_get_grid(dev, 0, out), "get_grid"); }
