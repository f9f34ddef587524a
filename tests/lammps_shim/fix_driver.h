// Generic C driver around a LAMMPS-style fix, compiled once per fix class.
// The harness (Python via ctypes) plays LAMMPS: it fills the per-atom arrays,
// hands over a full neighbour list and calls the hooks in Verlet order.  The
// same entry points exist for the compiled reference fix (prefix `ref_`) and
// for the product fix (prefix `b200_`), so parity tests issue identical call
// sequences against both.
//
// FixT must derive from LAMMPS_NS::Fix and provide the probe accessors used at
// the bottom (p_rho, p_w, p_xi, p_feph, p_frng, grid_size, grid_T).
#pragma once

#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "lammps_shim.h"

namespace shim_driver {

using namespace LAMMPS_NS;

template <class FixT>
struct World {
  LAMMPS lmp;
  std::unique_ptr<FixT> fix;
  NeighList list;
  std::vector<double> xs, vs, fs, mass;
  std::vector<double *> xr, vr, fr;
  std::vector<int> type, mask;
  std::vector<tagint> tag;
  std::vector<int> numneigh, flat;
  std::vector<int *> firstneigh;
  std::vector<double> xi;
  std::string err;
  int nmax_fix = 0;
};

template <class W, class F>
int guarded(W *w, F &&body) {
  try {
    body();
    return 0;
  } catch (const std::exception &e) {
    w->err = e.what();
    return -1;
  } catch (...) {
    w->err = "unknown exception";
    return -2;
  }
}

template <class FixT>
World<FixT> *world_new(long long natoms, int ntypes, const double *boxlo, const double *boxhi, double dt,
                       const double *mass_by_type) {
  auto *w = new World<FixT>;
  w->lmp.atom->natoms = natoms;
  w->lmp.atom->ntypes = ntypes;
  for (int d = 0; d < 3; ++d) {
    w->lmp.domain->boxlo[d] = boxlo[d];
    w->lmp.domain->boxhi[d] = boxhi[d];
  }
  w->lmp.update->dt = dt;
  w->lmp.comm->me = shim_mpi().rank;      // several ranks: the test has plugged its transport in before (P_set_mpi)
  w->lmp.comm->nprocs = shim_mpi().size;
  w->mass.assign(ntypes + 1, 1.0);
  for (int t = 1; t <= ntypes; ++t) w->mass[t] = mass_by_type ? mass_by_type[t - 1] : 1.0;
  w->lmp.atom->mass = w->mass.data();
  return w;
}

template <class FixT>
int world_set_atoms(World<FixT> *w, int nlocal, int nghost, const double *x, const double *v, const double *f,
                    const int *type, const int *mask, const long long *tag, const int *ghost_owner) {
  return guarded(w, [&] {
    size_t n = static_cast<size_t>(nlocal) + nghost;
    w->xs.assign(x, x + 3 * n);
    w->vs.assign(v, v + 3 * n);
    if (f) w->fs.assign(f, f + 3 * n); else w->fs.assign(3 * n, 0.0);
    w->type.assign(type, type + n);
    w->mask.assign(mask, mask + n);
    w->tag.resize(n);
    for (size_t i = 0; i < n; ++i) w->tag[i] = static_cast<tagint>(tag ? tag[i] : static_cast<long long>(i + 1));
    w->xr.resize(n); w->vr.resize(n); w->fr.resize(n);
    for (size_t i = 0; i < n; ++i) {
      w->xr[i] = &w->xs[3 * i]; w->vr[i] = &w->vs[3 * i]; w->fr[i] = &w->fs[3 * i];
    }
    Atom *a = w->lmp.atom;
    a->nlocal = nlocal; a->nghost = nghost; a->nmax = static_cast<int>(n);
    a->x = w->xr.data(); a->v = w->vr.data(); a->f = w->fr.data();
    a->type = w->type.data(); a->mask = w->mask.data(); a->tag = w->tag.data();
    w->lmp.comm->ghost_owner.assign(ghost_owner, ghost_owner + nghost);
    if (w->fix && a->nmax > w->nmax_fix) {   // LAMMPS' grow callback
      w->fix->grow_arrays(a->nmax);
      w->nmax_fix = a->nmax;
    }
  });
}

// Update positions / velocities / forces in place (same atom counts).
template <class FixT>
int world_update_xvf(World<FixT> *w, const double *x, const double *v, const double *f) {
  return guarded(w, [&] {
    if (x) std::copy(x, x + w->xs.size(), w->xs.begin());
    if (v) std::copy(v, v + w->vs.size(), w->vs.begin());
    if (f) std::copy(f, f + w->fs.size(), w->fs.begin());
  });
}

template <class FixT>
int world_make_fix(World<FixT> *w, int narg, const char **arg) {
  return guarded(w, [&] {
    std::vector<std::string> store(arg, arg + narg);
    std::vector<char *> argv;
    for (auto &s : store) argv.push_back(const_cast<char *>(s.c_str()));
    w->fix.reset(new FixT(&w->lmp, narg, argv.data()));
    w->nmax_fix = w->lmp.atom->nmax;
    w->fix->init();
    w->fix->init_list(0, &w->list);
  });
}

// Several ranks: the swaps of Comm::forward_comm(Fix*) (lammps_shim.h) -- per peer (this rank included, for its own
// periodic images) the local atoms it holds as ghosts and the contiguous ghost range it fills -- and the transport.
template <class FixT>
int world_set_swaps(World<FixT> *w, int nswaps, const int *peer, const int *send_count, const int *sendlist, const int *first,
                    const int *n, void (*exchange)(int, const int *, double *const *, const int *, double *const *, const int *)) {
  return guarded(w, [&] {
    Comm *c = w->lmp.comm;
    c->swaps.clear();
    size_t o = 0;
    for (int k = 0; k < nswaps; ++k) {
      Comm::Swap s;
      s.peer = peer[k];
      s.sendlist.assign(sendlist + o, sendlist + o + send_count[k]);
      o += send_count[k];
      s.first = first[k];
      s.n = n[k];
      c->swaps.push_back(s);
    }
    c->exchange = exchange;
  });
}

// CSR neighbour list: offsets[nlocal+1], flat[offsets[nlocal]]
template <class FixT>
int world_set_neighbors(World<FixT> *w, int nlocal, const long long *offsets, const int *flat) {
  return guarded(w, [&] {
    w->flat.assign(flat, flat + offsets[nlocal]);
    w->numneigh.resize(nlocal);
    w->firstneigh.resize(nlocal);
    for (int i = 0; i < nlocal; ++i) {
      w->numneigh[i] = static_cast<int>(offsets[i + 1] - offsets[i]);
      w->firstneigh[i] = w->flat.data() + offsets[i];
    }
    w->list.inum = nlocal;
    w->list.numneigh = w->numneigh.data();
    w->list.firstneigh = w->firstneigh.data();
    w->lmp.neighbor->ago = 0;
    if (w->fix) w->fix->init_list(0, &w->list);
  });
}

// xi for every local atom [nlocal][3]; replayed through the RanMars stand-in
// in the order the fix consumes it (group atoms, ascending local index).
template <class FixT>
int world_set_xi(World<FixT> *w, const double *xi) {
  return guarded(w, [&] {
    Atom *a = w->lmp.atom;
    w->xi.clear();
    for (int i = 0; i < a->nlocal; ++i)
      if (a->mask[i] & w->fix->groupbit)
        for (int d = 0; d < 3; ++d) w->xi.push_back(xi[3 * i + d]);
    RanMars::inject = w->xi.data();
    RanMars::inject_len = w->xi.size();
    RanMars::cursor = 0;
  });
}

// What LAMMPS' spatial sort (atom_modify sort, every 1000 steps by default) does to a fix: the local atoms are
// re-ordered -- the atom at index i moves to new_of_old[i] -- and the fix is told through copy_arrays(i, j, 0) calls
// that walk the cycles of the permutation with one scratch slot, exactly like AtomVec / Modify::copy_arrays in
// Atom::sort().  The harness' own arrays, the ghost->owner map and the neighbour list are relabelled accordingly, so
// the physics is unchanged and a fix that migrates its per-atom state correctly continues the same trajectory.
template <class FixT>
int world_permute(World<FixT> *w, const int *new_of_old) {
  return guarded(w, [&] {
    Atom *a = w->lmp.atom;
    const int nl = a->nlocal, nt = nl + a->nghost;
    std::vector<int> old_of_new(nl, -1);
    for (int i = 0; i < nl; ++i) {
      if (new_of_old[i] < 0 || new_of_old[i] >= nl || old_of_new[new_of_old[i]] != -1) throw std::runtime_error("permute: not a permutation");
      old_of_new[new_of_old[i]] = i;
    }
    if (w->nmax_fix < nt + 1) {   // one scratch slot behind the ghosts
      w->fix->grow_arrays(nt + 1);
      w->nmax_fix = nt + 1;
    }
    a->nmax = std::max(a->nmax, nt + 1);
    const int scratch = nt;
    std::vector<char> done(nl, 0);
    for (int j0 = 0; j0 < nl; ++j0) {
      if (done[j0] || old_of_new[j0] == j0) continue;
      w->fix->copy_arrays(j0, scratch, 0);
      int j = j0;
      for (;;) {
        done[j] = 1;
        const int src = old_of_new[j];
        if (src == j0) { w->fix->copy_arrays(scratch, j, 0); break; }
        w->fix->copy_arrays(src, j, 0);
        j = src;
      }
    }
    auto perm3 = [&](std::vector<double> &v) {
      std::vector<double> t(v.begin(), v.begin() + 3 * (size_t)nl);
      for (int i = 0; i < nl; ++i) for (int d = 0; d < 3; ++d) v[3 * (size_t)new_of_old[i] + d] = t[3 * (size_t)i + d];
    };
    perm3(w->xs); perm3(w->vs); perm3(w->fs);
    auto perm1 = [&](auto &v) {
      auto t = v;
      for (int i = 0; i < nl; ++i) v[new_of_old[i]] = t[i];
    };
    perm1(w->type); perm1(w->mask); perm1(w->tag);
    for (auto &o : w->lmp.comm->ghost_owner) o = new_of_old[o];
    // neighbour list: row j of the new order is the old row of old_of_new[j], entries relabelled
    std::vector<int> flat;
    std::vector<int> num(nl);
    flat.reserve(w->flat.size());
    std::vector<size_t> start(nl);
    for (int j = 0; j < nl; ++j) {
      const int i = old_of_new[j];
      start[j] = flat.size();
      num[j] = w->numneigh[i];
      for (int k = 0; k < w->numneigh[i]; ++k) {
        const int e = w->firstneigh[i][k];
        const int idx = e & NEIGHMASK, hi = e & ~NEIGHMASK;
        flat.push_back(hi | (idx < nl ? new_of_old[idx] : idx));
      }
    }
    w->flat.swap(flat);
    w->numneigh = num;
    for (int j = 0; j < nl; ++j) w->firstneigh[j] = w->flat.data() + start[j];
    w->list.numneigh = w->numneigh.data();
    w->list.firstneigh = w->firstneigh.data();
    w->lmp.neighbor->ago = 0;
    w->fix->init_list(0, &w->list);
  });
}

}  // namespace shim_driver

#define SHIM_DRIVER_DEFINE(P, FixT)                                                                            \
  const double *LAMMPS_NS::RanMars::inject = nullptr;                                                          \
  size_t LAMMPS_NS::RanMars::inject_len = 0;                                                                   \
  size_t LAMMPS_NS::RanMars::cursor = 0;                                                                       \
  extern "C" {                                                                                                 \
  typedef shim_driver::World<FixT> P##_world;                                                                  \
  void *P##_world_new(long long natoms, int ntypes, const double *lo, const double *hi, double dt,             \
                      const double *mass) {                                                                    \
    return shim_driver::world_new<FixT>(natoms, ntypes, lo, hi, dt, mass);                                     \
  }                                                                                                            \
  void P##_world_free(void *w) { delete static_cast<P##_world *>(w); }                                         \
  void P##_set_mpi(int rank, int size, void (*allreduce)(void *, int, int, int), void (*bcast)(void *, int, int),   \
                   void (*alltoallv)(const int *, const int *, const int *, int *, const int *, const int *),  \
                   void (*barrier)()) {                                                                        \
    ShimMpiBackend &b = shim_mpi();                                                                            \
    b.rank = rank; b.size = size; b.allreduce = allreduce; b.bcast = bcast; b.alltoallv_int = alltoallv;       \
    b.barrier = barrier;                                                                                       \
  }                                                                                                            \
  int P##_set_swaps(void *w, int nswaps, const int *peer, const int *send_count, const int *sendlist,          \
                    const int *first, const int *n,                                                            \
                    void (*exchange)(int, const int *, double *const *, const int *, double *const *, const int *)) { \
    return shim_driver::world_set_swaps(static_cast<P##_world *>(w), nswaps, peer, send_count, sendlist, first, n, \
                                        exchange);                                                             \
  }                                                                                                            \
  const char *P##_last_error(void *w) { return static_cast<P##_world *>(w)->err.c_str(); }                     \
  int P##_set_atoms(void *w, int nlocal, int nghost, const double *x, const double *v, const double *f,        \
                    const int *type, const int *mask, const long long *tag, const int *owner) {                \
    return shim_driver::world_set_atoms(static_cast<P##_world *>(w), nlocal, nghost, x, v, f, type, mask, tag, \
                                        owner);                                                                \
  }                                                                                                            \
  int P##_update_xvf(void *w, const double *x, const double *v, const double *f) {                             \
    return shim_driver::world_update_xvf(static_cast<P##_world *>(w), x, v, f);                                \
  }                                                                                                            \
  int P##_make_fix(void *w, int narg, const char **arg) {                                                      \
    return shim_driver::world_make_fix(static_cast<P##_world *>(w), narg, arg);                                \
  }                                                                                                            \
  int P##_set_neighbors(void *w, int nlocal, const long long *off, const int *flat) {                          \
    return shim_driver::world_set_neighbors(static_cast<P##_world *>(w), nlocal, off, flat);                   \
  }                                                                                                            \
  int P##_permute(void *w, const int *new_of_old) {                                                           \
    return shim_driver::world_permute(static_cast<P##_world *>(w), new_of_old);                               \
  }                                                                                                            \
  int P##_set_xi(void *w, const double *xi) {                                                                  \
    return shim_driver::world_set_xi(static_cast<P##_world *>(w), xi);                                         \
  }                                                                                                            \
  int P##_set_dt(void *w_, double dt) {                                                                        \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    return shim_driver::guarded(w, [&] { w->lmp.update->dt = dt; w->fix->reset_dt(); });                       \
  }                                                                                                            \
  void P##_set_step(void *w, long long step) { static_cast<P##_world *>(w)->lmp.update->ntimestep = step; }    \
  /* Neighbor::decide() of a step without re-neighbouring: the list ages */                                    \
  void P##_neigh_tick(void *w_) { ++static_cast<P##_world *>(w_)->lmp.neighbor->ago; }                         \
  void P##_neigh_modify(void *w_, int every, int delay, int check) {                                           \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    w->lmp.neighbor->every = every; w->lmp.neighbor->delay = delay; w->lmp.neighbor->dist_check = check;       \
  }                                                                                                            \
  int P##_initial_integrate(void *w_) {                                                                        \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    return shim_driver::guarded(w, [&] { w->fix->initial_integrate(0); });                                     \
  }                                                                                                            \
  int P##_post_force(void *w_) {                                                                               \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    return shim_driver::guarded(w, [&] { w->fix->post_force(0); });                                            \
  }                                                                                                            \
  int P##_final_integrate(void *w_) {                                                                          \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    return shim_driver::guarded(w, [&] { w->fix->final_integrate(); });                                        \
  }                                                                                                            \
  int P##_end_of_step(void *w_) {                                                                              \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    return shim_driver::guarded(w, [&] { w->fix->end_of_step(); });                                            \
  }                                                                                                            \
  int P##_post_run(void *w_) {                                                                                 \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    return shim_driver::guarded(w, [&] { w->fix->post_run(); });                                               \
  }                                                                                                            \
  int P##_setmask(void *w) { return static_cast<P##_world *>(w)->fix->setmask(); }                             \
  double P##_compute_vector(void *w, int i) { return static_cast<P##_world *>(w)->fix->compute_vector(i); }    \
  long long P##_n_forward(void *w) { return static_cast<P##_world *>(w)->lmp.comm->n_forward; }                \
  double P##_neigh_cutoff(void *w) { return static_cast<P##_world *>(w)->lmp.neighbor->last_request.cutoff; }  \
  int P##_fix_flags(void *w_, int *out) {                                                                      \
    auto *f = static_cast<P##_world *>(w_)->fix.get();                                                         \
    int v[] = {f->vector_flag, f->size_vector, f->global_freq, f->extvector, f->nevery, f->peratom_flag,       \
               f->size_peratom_cols, f->peratom_freq, f->comm_forward, f->time_integrate,                      \
               static_cast<P##_world *>(w_)->lmp.comm->ghost_velocity};                                        \
    for (int i = 0; i < 11; ++i) out[i] = v[i];                                                                \
    return 11;                                                                                                 \
  }                                                                                                            \
  void P##_get_xvf(void *w_, double *x, double *v, double *f) {                                                \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    if (x) std::copy(w->xs.begin(), w->xs.end(), x);                                                           \
    if (v) std::copy(w->vs.begin(), w->vs.end(), v);                                                           \
    if (f) std::copy(w->fs.begin(), w->fs.end(), f);                                                           \
  }                                                                                                            \
  void P##_get_array(void *w_, double *out) {                                                                  \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    int n = w->lmp.atom->nlocal, c = w->fix->size_peratom_cols;                                                \
    for (int i = 0; i < n; ++i)                                                                                \
      for (int k = 0; k < c; ++k) out[(size_t)i * c + k] = w->fix->array_atom[i][k];                           \
  }                                                                                                            \
  /* which: 0 rho[ntotal] 1 w[nlocal*3] 2 xi[nlocal*3] 3 f_EPH[nlocal*3] 4 f_RNG[nlocal*3] */                  \
  int P##_get_probe(void *w_, int which, double *out) {                                                        \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    return shim_driver::guarded(w, [&] {                                                                       \
      size_t nl = w->lmp.atom->nlocal, nt = nl + w->lmp.atom->nghost;                                          \
      w->fix->probe_copy(which, nl, nt, out);                                                                  \
    });                                                                                                        \
  }                                                                                                            \
  long long P##_grid_size(void *w) { return (long long)static_cast<P##_world *>(w)->fix->grid_size(); }        \
  int P##_grid_T(void *w_, double *out) {                                                                      \
    auto *w = static_cast<P##_world *>(w_);                                                                    \
    return shim_driver::guarded(w, [&] { w->fix->grid_T(out); });                                              \
  }                                                                                                            \
  }
