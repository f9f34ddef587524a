// Test-only stand-in for the slice of the LAMMPS core that a `fix eph`-style
// plug-in touches.  LAMMPS itself is third-party code that is neither vendored
// in the reference checkout nor installed in this image, so the harness plays
// its role: it owns the per-atom arrays, the full neighbour list, the box and
// the ghost->owner map, and it drives the fix hooks in Verlet order.
//
// The surface mirrored here is the one listed in SURVEY.md Appendix B (every
// LAMMPS symbol the reference fix_eph.cpp uses).  It is deliberately tiny:
// one rank, periodic ghost images, no MPI.  It is NOT part of the product.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <mpi.h>   // the serial stub that sits next to this header

#define FLERR __FILE__, __LINE__
#define NEIGHMASK 0x1FFFFFFF

namespace LAMMPS_NS {

// LAMMPS' integer widths depend on the build: -DLAMMPS_SMALLBIG (the default) has 32-bit atom tags and 64-bit atom counts
// and time steps, -DLAMMPS_BIGBIG 64-bit tags.  The stand-in follows BIGBIG unless built with -DSHIM_TAGINT32.
#ifdef SHIM_TAGINT32
typedef int tagint;
#else
typedef long long tagint;
#endif
typedef long long bigint;

class Fix;
class LAMMPS;

struct ShimError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class Error {
 public:
  [[noreturn]] void all(const char *file, int line, const std::string &msg) {
    throw ShimError(std::string(file) + ":" + std::to_string(line) + ": " + msg);
  }
  [[noreturn]] void one(const char *file, int line, const std::string &msg) { all(file, line, msg); }
  void warning(const char *, int, const std::string &msg) { std::fprintf(stderr, "WARNING: %s\n", msg.c_str()); }
};

class Atom {
 public:
  long long natoms = 0;
  int ntypes = 0;
  int nmax = 0;
  int nlocal = 0;
  int nghost = 0;
  double **x = nullptr, **v = nullptr, **f = nullptr;
  double *mass = nullptr;   // indexed by type, 1-based
  int *type = nullptr;
  int *mask = nullptr;
  tagint *tag = nullptr;
  int ncallbacks = 0;
  void add_callback(int) { ++ncallbacks; }
  void delete_callback(const char *, int) { --ncallbacks; }
};

class Domain {
 public:
  double boxlo[3] = {0, 0, 0};
  double boxhi[3] = {1, 1, 1};
  int dimension = 3;
  int nonperiodic = 0;
  int triclinic = 0;
};

class Update {
 public:
  double dt = 0.001;
  long long ntimestep = 0;
};

class Force {
 public:
  double boltz = 8.617343e-5;          // metal units, eV/K
  double ftm2v = 1.0 / 1.0364269e-4;   // metal units
};

class Memory {
 public:
  template <typename T>
  T *grow(T *&arr, int n, const char *) {
    arr = static_cast<T *>(std::realloc(arr, sizeof(T) * static_cast<size_t>(n > 0 ? n : 1)));
    return arr;
  }
  // contiguous 2-D array: row pointers into one slab (callers rely on &a[0][0])
  template <typename T>
  T **grow(T **&arr, int n1, int n2, const char *) {
    size_t rows = static_cast<size_t>(n1 > 0 ? n1 : 1);
    T *slab = arr ? arr[0] : nullptr;
    slab = static_cast<T *>(std::realloc(slab, sizeof(T) * rows * n2));
    arr = static_cast<T **>(std::realloc(arr, sizeof(T *) * rows));
    for (size_t i = 0; i < rows; ++i) arr[i] = slab + i * n2;
    return arr;
  }
  template <typename T>
  void destroy(T *&arr) {
    std::free(arr);
    arr = nullptr;
  }
  template <typename T>
  void destroy(T **&arr) {
    if (arr) {
      std::free(arr[0]);
      std::free(arr);
    }
    arr = nullptr;
  }
};

namespace NeighConst {
enum { REQ_DEFAULT = 0, REQ_FULL = 1 << 0, REQ_GHOST = 1 << 1 };
}

class NeighRequest {
 public:
  int style = 0;
  double cutoff = 0.0;
  void set_cutoff(double c) { cutoff = c; }
};

class NeighList {
 public:
  int inum = 0;
  int *ilist = nullptr;
  int *numneigh = nullptr;
  int **firstneigh = nullptr;
};

class Neighbor {
 public:
  NeighRequest last_request;
  double skin = 2.0;
  int ago = 0;
  int every = 1, delay = 0, dist_check = 1;   // neigh_modify every / delay / check (LAMMPS defaults: 1, 0, yes)
  NeighRequest *add_request(Fix *, int style) {
    last_request.style = style;
    return &last_request;
  }
};

class Comm {
 public:
  int ghost_velocity = 0;
  int me = 0;
  int nprocs = 1;
  // one rank: ghost g (array index nlocal+g) is an image of local atom ghost_owner[g]
  std::vector<int> ghost_owner;
  // several ranks: one swap per rank that holds atoms of this one as ghosts or owns ghosts of this one (this rank itself
  // included, for its own periodic images).  The ghosts a rank fills are one contiguous range of the atom arrays.
  struct Swap {
    int peer;
    std::vector<int> sendlist;   // local atoms the peer holds as ghosts, in the peer's ghost order
    int first, n;                // ghost range [first, first + n) the peer fills
  };
  std::vector<Swap> swaps;
  // moves the packed buffers of all swaps with other ranks at once (the test plugs torch.distributed in)
  void (*exchange)(int nswaps, const int *peer, double *const *sbuf, const int *sn, double *const *rbuf, const int *rn) = nullptr;
  Atom *atom = nullptr;
  long long n_forward = 0;
  inline void forward_comm(Fix *fix);
};

// Stand-in for LAMMPS' Marsaglia generator.  The real RanMars is third-party
// code absent from the reference checkout (RNG stream parity is "unpinned",
// SURVEY.md 8c); here gaussian() replays a stream injected by the harness so
// the reference fix and the product see identical xi.
class RanMars {
 public:
  RanMars(LAMMPS *, int seed) : seed_(seed) {}
  double gaussian() {
    if (!inject || cursor >= inject_len) throw ShimError("RanMars shim: injected stream exhausted");
    return inject[cursor++];
  }
  double uniform() { return 0.5; }
  static const double *inject;
  static size_t inject_len;
  static size_t cursor;

 private:
  int seed_;
};

class LAMMPS {
 public:
  Error *error;
  Atom *atom;
  Domain *domain;
  Neighbor *neighbor;
  Force *force;
  Update *update;
  Comm *comm;
  Memory *memory;
  MPI_Comm world = 0;
  LAMMPS() {
    error = new Error;
    atom = new Atom;
    domain = new Domain;
    neighbor = new Neighbor;
    force = new Force;
    update = new Update;
    comm = new Comm;
    memory = new Memory;
    comm->atom = atom;
  }
  ~LAMMPS() {
    delete error;
    delete atom;
    delete domain;
    delete neighbor;
    delete force;
    delete update;
    delete comm;
    delete memory;
  }
};

class Pointers {
 public:
  explicit Pointers(LAMMPS *ptr)
      : lmp(ptr), error(ptr->error), atom(ptr->atom), domain(ptr->domain), neighbor(ptr->neighbor),
        force(ptr->force), update(ptr->update), comm(ptr->comm), memory(ptr->memory), world(ptr->world) {}
  virtual ~Pointers() = default;

 protected:
  LAMMPS *lmp;
  Error *&error;
  Atom *&atom;
  Domain *&domain;
  Neighbor *&neighbor;
  Force *&force;
  Update *&update;
  Comm *&comm;
  Memory *&memory;
  MPI_Comm &world;
};

namespace FixConst {
enum {
  INITIAL_INTEGRATE = 1 << 0,
  POST_INTEGRATE = 1 << 1,
  PRE_EXCHANGE = 1 << 2,
  PRE_NEIGHBOR = 1 << 3,
  POST_NEIGHBOR = 1 << 4,
  PRE_FORCE = 1 << 5,
  PRE_REVERSE = 1 << 6,
  POST_FORCE = 1 << 7,
  FINAL_INTEGRATE = 1 << 8,
  END_OF_STEP = 1 << 9,
  POST_RUN = 1 << 10
};
}

class Fix : protected Pointers {
 public:
  char *id = nullptr;
  int igroup = 0, groupbit = 1;
  int vector_flag = 0, size_vector = 0, global_freq = 0, extvector = 0, nevery = 1;
  int peratom_flag = 0, size_peratom_cols = 0, peratom_freq = 0;
  int comm_forward = 0, time_integrate = 0;
  int maxexchange = 0;   // doubles per atom in pack_exchange (fix eph/coloured/exp carries its two filtered forces)
  double **array_atom = nullptr;
  double *vector_atom = nullptr;

  Fix(LAMMPS *l, int narg, char **arg) : Pointers(l) {
    if (narg > 0) {
      id = new char[std::strlen(arg[0]) + 1];
      std::strcpy(id, arg[0]);
    }
    // group "all" -> bit 1; any other group name "bitN" -> 1<<N (harness convention)
    if (narg > 1 && std::strncmp(arg[1], "bit", 3) == 0) groupbit = 1 << std::atoi(arg[1] + 3);
  }
  ~Fix() override { delete[] id; }

  virtual int setmask() = 0;
  virtual void init() {}
  virtual void init_list(int, NeighList *) {}
  virtual void setup(int) {}
  virtual void initial_integrate(int) {}
  virtual void post_force(int) {}
  virtual void final_integrate() {}
  virtual void end_of_step() {}
  virtual void post_run() {}
  virtual void reset_dt() {}
  virtual void grow_arrays(int) {}
  virtual double compute_vector(int) { return 0.0; }
  virtual double memory_usage() { return 0.0; }
  virtual int pack_forward_comm(int, int *, double *, int, int *) { return 0; }
  virtual void unpack_forward_comm(int, int, double *) {}
  // atom migration callbacks (fix eph/atomic carries its per-atom electronic energy along, fix_eph_atomic.cpp:939-955)
  virtual int pack_exchange(int, double *) { return 0; }
  virtual int unpack_exchange(int, double *) { return 0; }
  virtual void copy_arrays(int, int, int) {}
};

// One "swap": every ghost receives the value of its owner through the fix's own
// pack/unpack callbacks, exactly the path LAMMPS' Comm::forward_comm(Fix*) takes.
inline void Comm::forward_comm(Fix *fix) {
  ++n_forward;
  if (!swaps.empty()) {
    const int width = fix->comm_forward > 0 ? fix->comm_forward : 1;
    std::vector<std::vector<double>> sb(swaps.size()), rb(swaps.size());
    std::vector<int> peer, sn, rn;
    std::vector<double *> sp, rp;
    std::vector<int> per(swaps.size(), 0);
    for (size_t k = 0; k < swaps.size(); ++k) {
      Swap &w = swaps[k];
      sb[k].resize(w.sendlist.size() * width + 1);
      const int m = w.sendlist.empty() ? 0 : fix->pack_forward_comm((int)w.sendlist.size(), w.sendlist.data(), sb[k].data(), 0, nullptr);
      per[k] = w.sendlist.empty() ? 0 : m / (int)w.sendlist.size();
      if (w.peer == me) continue;
      peer.push_back(w.peer); sn.push_back(m); sp.push_back(sb[k].data());
    }
    // every rank packs the same state, so the doubles per atom agree between the two sides; a rank that sends nothing
    // to anybody learns the width from whoever sends to it by receiving at full width and using what arrives
    int wid = 0;
    for (int p : per) wid = p > wid ? p : wid;
    MPI_Allreduce(MPI_IN_PLACE, &wid, 1, MPI_INT, MPI_MAX, 0);
    for (size_t k = 0; k < swaps.size(); ++k) {
      Swap &w = swaps[k];
      rb[k].resize((size_t)w.n * wid + 1);
      if (w.peer == me) continue;
      rn.push_back(w.n * wid); rp.push_back(rb[k].data());
    }
    if (!peer.empty()) {
      if (!exchange) throw ShimError("Comm shim: several ranks but no transport plugged in");
      exchange((int)peer.size(), peer.data(), sp.data(), sn.data(), rp.data(), rn.data());
    }
    for (size_t k = 0; k < swaps.size(); ++k) {
      Swap &w = swaps[k];
      if (w.n == 0) continue;
      fix->unpack_forward_comm(w.n, w.first, w.peer == me ? sb[k].data() : rb[k].data());
    }
    return;
  }
  int n = static_cast<int>(ghost_owner.size());
  if (n == 0) return;
  std::vector<double> buf(static_cast<size_t>(n) * (fix->comm_forward > 0 ? fix->comm_forward : 1));
  fix->pack_forward_comm(n, ghost_owner.data(), buf.data(), 0, nullptr);
  fix->unpack_forward_comm(n, atom->nlocal, buf.data());
}

}  // namespace LAMMPS_NS
