// C ABI of the B200-native `fix eph` hot path (see include/eph_b200.h).
// Host-side handle, buffer management and kernel orchestration.  There is no
// CPU fallback: every entry point either enqueues CUDA work or fails.
#include "eph_b200.h"

#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "eph_atoms.cuh"
#include "eph_device.cuh"
#include "eph_grid.cuh"
#include "eph_grid_tma.cuh"
#include "eph_neigh.cuh"
#include "eph_sweeps.cuh"
#include "eph_legacy.cuh"
#include "eph_nccl.h"
#include "eph_p2p.cuh"

using namespace ephb;

namespace {

std::string g_create_error;

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 512;   // slack: the list streams of the sweeps are prefetched a few iterations past the end
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct eph_b200_handle {
  eph_b200_config cfg{};
  std::vector<int> type_map;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int max_smem_optin = 0;
  std::string err;
  long long launches = 0;

  // optional per-kernel timing with CUDA events on the launch stream
  int profiling = 0;   // 0 off, 1 every kernel, 2 the two list sweeps only
  struct KernelStat { const char *name; double ms = 0; long long count = 0; };
  std::vector<KernelStat> kstats;
  struct Pending { int stat; cudaEvent_t beg, end; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> event_pool;

  // tables
  int n_el = 0, n_rho = 0, n_beta = 0;
  double inv_dr_sq = 0, inv_drho = 0, rc2 = 0, rho_cut = 0;
  DevBuf<double2> rho_tab, alpha_tab, beta_tab;
  DevBuf<double2> rho_r_tab;   // rho(r) splines in r: eph_model 2 only
  double inv_dr = 0;
  bool rho_r_set = false;
  DevBuf<int> d_type_map;
  bool tables_set = false;

  // time step
  double dt = 0, boltz = 0, eta = 0;
  bool dt_set = false;

  // atoms (LAMMPS order)
  int nlocal = 0, nghost = 0;
  bool atoms_set = false;
  DevBuf<int> type, mask, owner;
  DevBuf<long long> tag;
  bool has_owner = false;
  DevBuf<double> x, v, f, xi_in;  // staging for host memspace
  DevBuf<int> comm_idx;           // forward-comm scratch (host transport)
  DevBuf<double> comm_buf;
  DevBuf<double> mass;
  std::vector<double> mass_host;   // what h->mass holds: the masses go up again only when they change

  // internal per-atom records
  DevBuf<double4> pos4, pv, puz, W4;
  DevBuf<double> rho, w, xi, f_eph, f_rng, array8, gpair, gpair_i;
  // packed gather records (eph_packed.cuh): half the sectors per list slot of the fp64 records pv / puz
  DevBuf<Packed32> recD, recA;
  DevBuf<Block16> recB;
  DevBuf<double> var;
  int precision = 1;            // 1: packed records where they apply (default), 0: fp64 records everywhere
  bool step_packed = false;     // this step's density pass ran on packed records (the fp64 pass stands by as fall-back)
  // `fix eph/coloured/exp` (fix_eph_coloured_exp.cpp): exponential memory kernel on both forces; the filtered forces of
  // the previous step are per-atom state (f_dis, f_sto [nlocal][3])
  bool coloured = false;
  double tau0 = 0.0, zeta = 0.0;
  DevBuf<double> f_dis, f_sto;
  bool forces_valid = false;
  bool peratom_valid = false;   // an end_of_step has run: the per-atom output (fix_eph.cpp:406-428) can be materialised
  // state between post_force_begin and post_force_end
  bool pf_open = false, pf_build = false;
  const double *pf_xi = nullptr;
  long long pf_step = 0;
  bool eos_open = false;
  // optional second stream for the grid: the source all-reduce and the solve overlap the next step's density pass
  cudaStream_t grid_stream = nullptr;
  cudaEvent_t ev_deposit = nullptr, ev_solved = nullptr;
  bool solve_pending = false;
  // optional communication stream + boundary-first density pass: the ghost exchange of a step (pack, the caller's
  // all-to-all, unpack) runs on it behind the boundary tiles while the main stream still sweeps the interior tiles
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_boundary = nullptr, ev_unpacked = nullptr;
  bool unpack_pending = false, boundary_recorded = false;
  DevBuf<int> tile_flag, tile_scan, work_all;   // work_all: boundary tiles, then interior tiles
  int n_first = 0, n_rest = 0;          // boundary / interior tiles
  DevBuf<unsigned> done_counter;        // finished boundary tiles (monotonic, wraps; compared cyclically)
  unsigned boundary_target = 0;         // counter value at which this step's boundary tiles are all done
  bool boundary_by_counter = false;     // this step's pack waits on the counter (one launch) instead of an event (two)
  bool split_ready = false;
  double *dT_e_ext = nullptr;   // caller-owned grid source term (multi-rank: all-reduced between the two end_of_step halves)

  // multi-rank data plane inside the engine (NCCL over NVLink; eph_b200_comm_init / set_ghost_map): the ghost exchange
  // of post_force, the all-reduce of the grid source term and the halo planes of the sharded grid solve
  NcclComm comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  bool ghost_map_set = false;
  std::vector<int> gm_peer, gm_send_count, gm_recv_count;   // per peer rank
  DevBuf<int> gm_send_idx, gm_recv_slot;                    // concatenated over the peers, in peer order
  int gm_nsend = 0, gm_nrecv = 0;
  DevBuf<double> gm_send_buf, gm_recv_buf, gm_send_xi, gm_recv_xi;   // {rho, Wx, Wy, Wz} per atom; xi only when injected
  // ghost positions and velocities following their owners on the device (eph_b200_refresh_ghosts): shifts recorded at
  // the first call after set_atoms, when the caller's ghost coordinates are LAMMPS' own
  DevBuf<double> gshift;        // [nghost][3] x_ghost - x_owner
  bool gshift_valid = false;
  DevBuf<double> gm_send_xv, gm_recv_xv;   // {x, v} of the atoms other ranks hold as ghosts
  // the same two exchanges through NVLink peer memory (eph_p2p.cuh) when every rank of the communicator can map every
  // other rank's window (one node); NCCL send/receive otherwise
  bool p2p_ok = false;
  char *p2p_window = nullptr;
  size_t p2p_window_bytes = 0, p2p_region_bytes = 0;
  std::vector<char *> p2p_peer_window;     // by rank; own entry = p2p_window
  P2PMap p2p_map{};
  unsigned long long p2p_epoch[2] = {0, 0};
  DevBuf<unsigned> p2p_done;               // finished blocks of the sending kernels, per kind
  DevBuf<double> p2p_scratch;
  // device-resident integration (eph_b200_resident_*): x, v of all atoms and f of the local ones stay here between hooks
  DevBuf<double> res_x, res_v, res_f;
  DevBuf<double> res_fe;        // this step's f_EPH + f_RNG, formed while the pair forces are still on their way up
  bool resident = false, res_f_valid = false, res_pf_started = false;
  bool grid_sharded = false;    // every rank advances only its z-slab of the grid (halo planes + all-gather)
  DevBuf<double> slab_tmp;

  // neighbours
  DevBuf<long long> off;
  DevBuf<int> neigh;
  long long n_entries = 0;
  bool neigh_set = false;
  const long long *off_ptr = nullptr;  // device pointers actually used (own or caller's)
  const int *neigh_ptr = nullptr;

  // device-side neighbour list construction (eph_b200_build_neighbors)
  DevBuf<int> nb_cell, nb_atom, nb_cell_s, nb_atom_s, nb_start, nb_end;
  DevBuf<double4> nb_xs;
  DevBuf<long long> nb_counts;
  DevBuf<unsigned char> nb_tmp;
  DevBuf<double> nb_box;

  // two-level Verlet list: inner list with a small skin, rebuilt on the device from LAMMPS' list
  DevBuf<int> ineigh, icount;
  DevBuf<long long> tile_caps, tile_off;  // warp-tiled layout of the inner list (eph_sweeps.cuh)
  int lanes = 2;                          // lanes per atom in both sweeps (fixes the tile shape); 2 measured best with packed records
  DevBuf<double4> xref, xref0;          // positions at the last inner build / at LAMMPS' build
  DevBuf<ListState> lstate;
  double skin = -1.0;                    // LAMMPS' neighbor->skin (unknown: inner list only valid from LAMMPS' build)
  double inner_skin = 0.4;
  bool inner_enabled = true;
  bool inner_wanted = true;              // the setting the next set_neighbors applies (set_skin with a list registered)
  bool fresh_neighbors = false;          // set_neighbors since the last post_force
  bool inner_prebuilt = false;           // eph_b200_build_neighbors filled the inner list's tiles while it wrote the full list
  bool have_inner = false;               // an inner list (and xref) exists
  bool inner_gave_up = false;            // rebuild was refused by the device-side check: use LAMMPS' list until it changes
  bool rebuilt_last_step = false;
  unsigned *h_flag = nullptr;            // pinned mirror of lstate.inner_invalid
  cudaEvent_t flag_event = nullptr;
  bool flag_pending = false;
  cudaStream_t copy_stream = nullptr;    // host memspace: f goes up while the density pass runs
  cudaEvent_t f_event = nullptr;
  bool f_prefetched = false;
  long long inner_builds = 0, inner_fallback_steps = 0;

  // grid
  int nx = 0, ny = 0, nz = 0, steps = 1;
  long long ncell = 0;
  double box[6] = {0, 1, 0, 1, 0, 1};
  double gdx = 1, gdy = 1, gdz = 1, dV = 1;
  DevBuf<double> T[2], dT_e, S_e, rho_e, C_e, kappa_e;
  int cur = 0;
  DevBuf<short> flag;
  DevBuf<unsigned short> t_dyn;
  bool grid_set = false, has_tdyn = false, minmax_valid = false;
  double c_min = 0, rho_min = 0, kappa_max = 0;
  int n_T = 0;
  double dT_tab = 0;
  DevBuf<double2> C_T_tab, K_T_tab;
  DevBuf<double> E_T_tab;
  int last_substeps = 0;
  double plan_inner_dt = 0;              // sub-step length of the last grid_plan
  bool plan_open = false;                // grid_plan_substeps called, sub-steps of this solve still to come
  int plan_done = 0;                     // sub-steps of the open plan already run
  // TMA path of the stencil: tensor maps over T_e (both buffers) and kappa_e; needs 16-byte row strides (nx even)
  bool tma_ok = false, has_walls = false;
  CUtensorMap map_T[2], map_K, map_S;
  // constant-coefficient fast path: all cells dynamic with identical parameters
  bool uniform = false;
  double u_kappa = 0, u_S = 0, u_rho = 0, u_C = 0;
  double *map_S_base = nullptr;

  // scalars
  DevBuf<double> d_scal;              // [0] E_local, [1] T sum
  DevBuf<unsigned long long> d_mm;    // min/max bits
  DevBuf<unsigned> d_status;
  double *h_pinned = nullptr;         // 8 doubles
};

namespace {

int fail(eph_b200_handle *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

int stat_index(eph_b200_handle *h, const char *name) {
  for (size_t i = 0; i < h->kstats.size(); ++i)
    if (std::strcmp(h->kstats[i].name, name) == 0) return (int)i;
  eph_b200_handle::KernelStat st;
  st.name = name;
  h->kstats.push_back(st);
  return (int)h->kstats.size() - 1;
}

cudaEvent_t take_event(eph_b200_handle *h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// RAII bracket around one kernel launch: two events on the launch stream when profiling is on
struct KernelTimer {
  eph_b200_handle *h;
  int idx = -1;
  cudaEvent_t beg = nullptr;
  cudaStream_t st;
  KernelTimer(eph_b200_handle *h_, const char *name, cudaStream_t st_ = nullptr) : h(h_), st(st_ ? st_ : h_->stream) {
    if (!h->profiling) return;
    if (h->profiling == 2 && std::strcmp(name, "density_sweep") != 0 && std::strcmp(name, "force_sweep") != 0) return;
    idx = stat_index(h, name);
    beg = take_event(h);
    cudaEventRecord(beg, st);
  }
  ~KernelTimer() {
    if (idx < 0) return;
    cudaEvent_t end = take_event(h);
    cudaEventRecord(end, st);
    h->pending.push_back({idx, beg, end});
  }
};

void drain_timers(eph_b200_handle *h) {
  for (auto &p : h->pending) {
    cudaEventSynchronize(p.end);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.beg, p.end) == cudaSuccess) {
      h->kstats[p.stat].ms += ms;
      h->kstats[p.stat].count += 1;
    }
    h->event_pool.push_back(p.beg);
    h->event_pool.push_back(p.end);
  }
  h->pending.clear();
}

#define EPH_CUDA(h, call)                                                                                \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(h, EPH_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define EPH_LAUNCH_CHECK(h)                                                                              \
  do {                                                                                                   \
    ++(h)->launches;                                                                                     \
    cudaError_t e_ = cudaGetLastError();                                                                 \
    if (e_ != cudaSuccess)                                                                               \
      return fail(h, EPH_B200_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// unmaps the peers' windows and frees the own one (eph_p2p.cuh)
void p2p_teardown(eph_b200_handle *h) {
  for (size_t r = 0; r < h->p2p_peer_window.size(); ++r)
    if (h->p2p_peer_window[r] && h->p2p_peer_window[r] != h->p2p_window) cudaIpcCloseMemHandle(h->p2p_peer_window[r]);
  h->p2p_peer_window.clear();
  if (h->p2p_window) cudaFree(h->p2p_window);
  h->p2p_window = nullptr;
  h->p2p_ok = false;
  h->p2p_done.release(); h->p2p_scratch.release();
  cudaGetLastError();
}

// masses by type on the device; uploaded only when they differ from what is there (a copy per hook otherwise)
int upload_masses(eph_b200_handle *h, const double *mass_by_type) {
  const size_t n = (size_t)h->cfg.ntypes + 1;
  if (h->mass_host.size() == n && std::memcmp(h->mass_host.data(), mass_by_type, n * sizeof(double)) == 0) return EPH_B200_OK;
  EPH_CUDA(h, h->mass.reserve(n));
  h->mass_host.assign(mass_by_type, mass_by_type + n);
  // the vector outlives the copy: a pageable source is staged before cudaMemcpyAsync returns anyway
  EPH_CUDA(h, cudaMemcpyAsync(h->mass.p, h->mass_host.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return EPH_B200_OK;
}

// the main stream must not read T_e / write dT_e before a solve running on the grid stream has finished
inline void join_grid_stream(eph_b200_handle *h) {
  if (h->solve_pending) {
    cudaStreamWaitEvent(h->stream, h->ev_solved, 0);
    h->solve_pending = false;
  }
}

inline int blocks_for(long long n, int threads) { return (int)std::max<long long>(1, (n + threads - 1) / threads); }

GridGeom grid_geom(const eph_b200_handle *h) {
  GridGeom g;
  g.nx = h->nx; g.ny = h->ny; g.nz = h->nz;
  g.x0 = h->box[0]; g.y0 = h->box[2]; g.z0 = h->box[4];
  g.dx = h->gdx; g.dy = h->gdy; g.dz = h->gdz;
  return g;
}

// copy-or-alias: host memspace stages into `buf`, device memspace uses the caller's pointer
template <class T>
int stage_in(eph_b200_handle *h, DevBuf<T> &buf, const T *src, size_t n, int memspace, const T **out) {
  if (memspace == EPH_B200_DEVICE) {
    *out = src;
    return EPH_B200_OK;
  }
  EPH_CUDA(h, buf.reserve(n));
  EPH_CUDA(h, cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  *out = buf.p;
  return EPH_B200_OK;
}

// cuStreamWaitValue32 through the runtime's driver entry point query (no -lcuda at link time); nullptr if unavailable
typedef CUresult (*StreamWaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValueFn stream_wait_value() {
  static StreamWaitValueFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char *off = std::getenv("EPH_B200_NO_WAIT_VALUE");
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (!(off && off[0] == '1') && cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && p)
      fn = reinterpret_cast<StreamWaitValueFn>(p);
  }
  return fn;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D fp64 tensor map over an [nz][ny][nx] field with a (kBX, kBY, kBZ) box; out-of-range elements read as zero
bool encode_grid_map(CUtensorMap *map, double *base, int nx, int ny, int nz, int bx = kBX, int by = kBY, int bz = kBZ) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
  cuuint64_t strides[2] = {(cuuint64_t)nx * sizeof(double), (cuuint64_t)nx * ny * sizeof(double)};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- packed records (eph_packed.cuh) ----
// Period of the packed positions: a power of two comfortably above twice the largest pair distance a walk of the inner
// list can meet (r_c + inner_skin when the list is built, plus inner_skin of relative drift before the displacement
// guard of pack_atoms trips).
double packed_period(const eph_b200_handle *h) {
  const double reach = std::sqrt(h->rc2) + 2.0 * h->inner_skin;
  double P = 1.0;
  while (P < 2.2 * reach) P *= 2.0;
  return P;
}
// Packed records serve model PRL with at most four elements (two flag bits) and need the inner list; a period above
// 64 A would make the position quantum too coarse for the 1e-10 bar.
bool packed_possible(const eph_b200_handle *h) {
  return h->precision == 1 && h->tables_set && h->cfg.model == EPH_B200_MODEL_PRL && h->n_el <= 4 && h->inner_enabled &&
         packed_period(h) <= 64.0;
}
}  // namespace

static int prepare_tiles(eph_b200_handle *h, int nlocal, long long total, const long long *total_pending);
static int register_list(eph_b200_handle *h, int nlocal, long long total, bool tiles_ready, bool inner_filled);

extern "C" {

int eph_b200_version(void) { return EPH_B200_VERSION; }
int eph_b200_device_count(int *out) {
  if (!out) return EPH_B200_ERR_ARG;
  *out = 0;
  return cudaGetDeviceCount(out) == cudaSuccess ? EPH_B200_OK : EPH_B200_ERR_NODEVICE;
}
const char *eph_b200_create_error(void) { return g_create_error.c_str(); }
const char *eph_b200_last_error(const eph_b200_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }
long long eph_b200_launch_count(const eph_b200_handle *h) { return h ? h->launches : 0; }

int eph_b200_create(const eph_b200_config *cfg, eph_b200_handle **out) {
  if (!cfg || !out) { g_create_error = "eph_b200_create: null argument"; return EPH_B200_ERR_ARG; }
  *out = nullptr;
  if (cfg->ntypes < 1 || !cfg->type_map) { g_create_error = "eph_b200_create: ntypes < 1 or no type_map"; return EPH_B200_ERR_ARG; }
  if (cfg->model != EPH_B200_MODEL_PRL && cfg->model != EPH_B200_MODEL_NONE && cfg->model != EPH_B200_MODEL_TTM &&
      cfg->model != EPH_B200_MODEL_PRB) {
    g_create_error = "eph_b200_create: eph_model must be 4 (PRL 120, 185501), 1 (TTM), 2 (PRB) or 0; model 3 reads out of "
                     "bounds in the reference (fix_eph.cpp:601) and is not offered";
    return EPH_B200_ERR_MODEL;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1) {
    g_create_error = std::string("eph_b200_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    return EPH_B200_ERR_NODEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "eph_b200_create: bad device ordinal"; return EPH_B200_ERR_ARG; }
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) {
    g_create_error = std::string("eph_b200_create: ") + cudaGetErrorString(e);
    return EPH_B200_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "eph_b200_create: kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return EPH_B200_ERR_NODEVICE;
  }
  auto *h = new eph_b200_handle;
  h->cfg = *cfg;
  h->type_map.assign(cfg->type_map, cfg->type_map + cfg->ntypes);
  h->cfg.type_map = h->type_map.data();
  if (const char *e = std::getenv("EPH_B200_INNER_SKIN")) {
    h->inner_skin = std::atof(e);
    h->inner_enabled = h->inner_wanted = h->inner_skin > 0.0;
  }
  if (const char *e = std::getenv("EPH_B200_RECORDS")) {   // gather records: "exact" (fp64) or "packed" (default)
    if (std::strcmp(e, "exact") == 0) h->precision = 0;
    else if (std::strcmp(e, "packed") == 0) h->precision = 1;
  }
  if (const char *e = std::getenv("EPH_B200_LANES")) {   // tuning knob: lanes per atom in the sweeps
    const int v = std::atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) h->lanes = v;
  }
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (cfg->stream) h->stream = static_cast<cudaStream_t>(cfg->stream);
  else {
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
      g_create_error = std::string("eph_b200_create: ") + cudaGetErrorString(e);
      delete h;
      return EPH_B200_ERR_CUDA;
    }
    h->own_stream = true;
  }
  bool ok = h->d_type_map.reserve(cfg->ntypes) == cudaSuccess && h->d_scal.reserve(8) == cudaSuccess &&
            h->lstate.reserve(1) == cudaSuccess && cudaMemset(h->lstate.p, 0, sizeof(ListState)) == cudaSuccess &&
            cudaMallocHost(&h->h_flag, sizeof(unsigned)) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->flag_event, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->f_event, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_deposit, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_solved, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_boundary, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_unpacked, cudaEventDisableTiming) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            h->d_mm.reserve(4) == cudaSuccess && h->d_status.reserve(1) == cudaSuccess &&
            cudaMallocHost(&h->h_pinned, 8 * sizeof(double)) == cudaSuccess &&
            cudaMemcpy(h->d_type_map.p, h->type_map.data(), cfg->ntypes * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemset(h->d_status.p, 0, sizeof(unsigned)) == cudaSuccess &&
            cudaMemset(h->d_scal.p, 0, 8 * sizeof(double)) == cudaSuccess;
  if (!ok) {
    g_create_error = std::string("eph_b200_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    eph_b200_destroy(h);
    return EPH_B200_ERR_CUDA;
  }
  *out = h;
  return EPH_B200_OK;
}

int eph_b200_destroy(eph_b200_handle *h) {
  if (!h) return EPH_B200_OK;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  h->rho_tab.release(); h->rho_r_tab.release(); h->alpha_tab.release(); h->beta_tab.release(); h->d_type_map.release();
  h->type.release(); h->mask.release(); h->owner.release(); h->tag.release();
  h->x.release(); h->v.release(); h->f.release(); h->xi_in.release(); h->mass.release(); h->comm_idx.release(); h->comm_buf.release();
  h->pos4.release(); h->pv.release(); h->puz.release(); h->W4.release(); h->gpair.release(); h->gpair_i.release();
  h->rho.release(); h->w.release(); h->xi.release(); h->f_eph.release(); h->f_rng.release(); h->array8.release();
  h->f_dis.release(); h->f_sto.release();
  h->recD.release(); h->recA.release(); h->recB.release(); h->var.release();
  p2p_teardown(h);
  if (h->comm) { nccl_api().CommDestroy(h->comm); h->comm = nullptr; }
  h->gm_send_idx.release(); h->gm_recv_slot.release(); h->gm_send_buf.release(); h->gm_recv_buf.release();
  h->gm_send_xi.release(); h->gm_recv_xi.release(); h->slab_tmp.release();
  h->gshift.release(); h->gm_send_xv.release(); h->gm_recv_xv.release(); h->res_x.release(); h->res_v.release(); h->res_f.release(); h->res_fe.release();
  h->off.release(); h->neigh.release();
  h->T[0].release(); h->T[1].release(); h->dT_e.release(); h->S_e.release(); h->rho_e.release(); h->C_e.release();
  h->kappa_e.release(); h->flag.release(); h->t_dyn.release(); h->C_T_tab.release(); h->K_T_tab.release(); h->E_T_tab.release();
  h->d_scal.release(); h->d_mm.release(); h->d_status.release();
  drain_timers(h);
  for (auto e : h->event_pool) cudaEventDestroy(e);
  h->nb_cell.release(); h->nb_atom.release(); h->nb_cell_s.release(); h->nb_atom_s.release(); h->nb_start.release();
  h->nb_end.release(); h->nb_xs.release(); h->nb_counts.release(); h->nb_tmp.release(); h->nb_box.release();
  h->ineigh.release(); h->icount.release(); h->tile_caps.release(); h->tile_off.release(); h->xref.release(); h->xref0.release(); h->lstate.release();
  if (h->h_flag) cudaFreeHost(h->h_flag);
  if (h->flag_event) cudaEventDestroy(h->flag_event);
  if (h->f_event) cudaEventDestroy(h->f_event);
  if (h->ev_deposit) cudaEventDestroy(h->ev_deposit);
  if (h->ev_solved) cudaEventDestroy(h->ev_solved);
  if (h->ev_boundary) cudaEventDestroy(h->ev_boundary);
  if (h->ev_unpacked) cudaEventDestroy(h->ev_unpacked);
  h->tile_flag.release(); h->tile_scan.release(); h->work_all.release(); h->done_counter.release();
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return EPH_B200_OK;
}

int eph_b200_set_profiling(eph_b200_handle *h, int on) {
  if (!h) return EPH_B200_ERR_ARG;
  drain_timers(h);
  h->profiling = on < 0 ? 0 : (on > 2 ? 1 : on);
  if (on) for (auto &k : h->kstats) { k.ms = 0; k.count = 0; }
  return EPH_B200_OK;
}

int eph_b200_kernel_times(eph_b200_handle *h, int max, const char **names, double *ms, long long *counts) {
  if (!h) return EPH_B200_ERR_ARG;
  cudaSetDevice(h->cfg.device);
  drain_timers(h);
  int n = 0;
  for (auto &k : h->kstats) {
    if (n >= max) break;
    if (names) names[n] = k.name;
    if (ms) ms[n] = k.ms;
    if (counts) counts[n] = k.count;
    ++n;
  }
  return n;
}

int eph_b200_synchronize(eph_b200_handle *h) {
  if (!h) return EPH_B200_ERR_ARG;
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->grid_stream) EPH_CUDA(h, cudaStreamSynchronize(h->grid_stream));
  return EPH_B200_OK;
}

int eph_b200_status_word(eph_b200_handle *h, unsigned *out) {
  if (!h || !out) return EPH_B200_ERR_ARG;
  EPH_CUDA(h, cudaMemcpyAsync(out, h->d_status.p, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPH_B200_OK;
}

int eph_b200_set_tables(eph_b200_handle *h, int n_elements, int n_rho, double inv_dr_sq, const double *coeff_rho_r_sq,
                        int n_beta, double inv_drho, const double *coeff_alpha, const double *coeff_beta,
                        double r_cutoff_sq, double rho_cutoff) {
  if (!h) return EPH_B200_ERR_ARG;
  if (n_elements < 1 || n_elements > 255 || n_rho < 4 || n_beta < 4 || !coeff_rho_r_sq || !coeff_alpha || !coeff_beta)
    return fail(h, EPH_B200_ERR_ARG, "set_tables: bad table sizes or null coefficients");
  for (int t = 0; t < h->cfg.ntypes; ++t)
    if (h->type_map[t] < 0 || h->type_map[t] >= n_elements)
      return fail(h, EPH_B200_ERR_ARG, "set_tables: type %d maps to element %d, table has %d", t + 1, h->type_map[t], n_elements);
  cudaSetDevice(h->cfg.device);
  size_t nr = (size_t)n_elements * n_rho * 2, nb = (size_t)n_elements * n_beta * 2;
  EPH_CUDA(h, h->rho_tab.reserve(nr));
  EPH_CUDA(h, h->alpha_tab.reserve(nb));
  EPH_CUDA(h, h->beta_tab.reserve(nb));
  EPH_CUDA(h, cudaMemcpyAsync(h->rho_tab.p, coeff_rho_r_sq, nr * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->alpha_tab.p, coeff_alpha, nb * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->beta_tab.p, coeff_beta, nb * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_el = n_elements; h->n_rho = n_rho; h->n_beta = n_beta;
  h->inv_dr_sq = inv_dr_sq; h->inv_drho = inv_drho; h->rc2 = r_cutoff_sq; h->rho_cut = rho_cutoff;
  h->tables_set = true;
  return EPH_B200_OK;
}

int eph_b200_set_rho_r_table(eph_b200_handle *h, int n_elements, int n_rho, double inv_dr, const double *coeff_rho_r) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->tables_set) return fail(h, EPH_B200_ERR_ARG, "set_rho_r_table: call set_tables first");
  if (n_elements != h->n_el || n_rho != h->n_rho || !coeff_rho_r || !(inv_dr > 0.0))
    return fail(h, EPH_B200_ERR_ARG, "set_rho_r_table: sizes differ from set_tables or null coefficients");
  if (h->cfg.ntypes > n_elements)   // the reference indexes rho(r) with type - 1 (fix_eph.cpp:530)
    return fail(h, EPH_B200_ERR_ARG, "set_rho_r_table: model 2 indexes rho(r) with the atom type: %d types, %d elements", h->cfg.ntypes, n_elements);
  cudaSetDevice(h->cfg.device);
  const size_t nr = (size_t)n_elements * n_rho * 2;
  EPH_CUDA(h, h->rho_r_tab.reserve(nr));
  EPH_CUDA(h, cudaMemcpyAsync(h->rho_r_tab.p, coeff_rho_r, nr * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  h->inv_dr = inv_dr;
  h->rho_r_set = true;
  return EPH_B200_OK;
}

int eph_b200_set_grid(eph_b200_handle *h, int nx, int ny, int nz, const double *box, int steps, const double *T_e,
                      const double *S_e, const double *rho_e, const double *C_e, const double *kappa_e,
                      const int16_t *flag, const uint16_t *t_dyn) {
  if (!h) return EPH_B200_ERR_ARG;
  if (nx < 1 || ny < 1 || nz < 1) return fail(h, EPH_B200_ERR_ARG, "FixEPH: non-positive grid values");
  if (!box || !T_e || !rho_e || !C_e || !kappa_e) return fail(h, EPH_B200_ERR_ARG, "set_grid: null field");
  if (!(box[0] < box[1] && box[2] < box[3] && box[4] < box[5])) return fail(h, EPH_B200_ERR_ARG, "set_grid: empty box");
  cudaSetDevice(h->cfg.device);
  long long n = (long long)nx * ny * nz;
  h->nx = nx; h->ny = ny; h->nz = nz; h->ncell = n; h->steps = steps > 0 ? steps : 1;
  std::memcpy(h->box, box, sizeof h->box);
  h->gdx = (box[1] - box[0]) / nx;  // eph_fdm.h:135-139
  h->gdy = (box[3] - box[2]) / ny;
  h->gdz = (box[5] - box[4]) / nz;
  h->dV = h->gdx * h->gdy * h->gdz;
  EPH_CUDA(h, h->T[0].reserve(n)); EPH_CUDA(h, h->T[1].reserve(n)); EPH_CUDA(h, h->dT_e.reserve(n));
  EPH_CUDA(h, h->S_e.reserve(n)); EPH_CUDA(h, h->rho_e.reserve(n)); EPH_CUDA(h, h->C_e.reserve(n));
  EPH_CUDA(h, h->kappa_e.reserve(n)); EPH_CUDA(h, h->flag.reserve(n)); EPH_CUDA(h, h->t_dyn.reserve(n));
  h->cur = 0;
  const size_t bytes = n * sizeof(double);
  EPH_CUDA(h, cudaMemcpyAsync(h->T[0].p, T_e, bytes, cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaMemsetAsync(h->dT_e.p, 0, bytes, h->stream));
  if (S_e) EPH_CUDA(h, cudaMemcpyAsync(h->S_e.p, S_e, bytes, cudaMemcpyHostToDevice, h->stream));
  else EPH_CUDA(h, cudaMemsetAsync(h->S_e.p, 0, bytes, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->rho_e.p, rho_e, bytes, cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->C_e.p, C_e, bytes, cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->kappa_e.p, kappa_e, bytes, cudaMemcpyHostToDevice, h->stream));
  std::vector<short> fl(n, 1);
  std::vector<unsigned short> td(n, 0);
  if (flag) std::copy(flag, flag + n, fl.begin());
  h->has_tdyn = false;
  if (t_dyn) {
    std::copy(t_dyn, t_dyn + n, td.begin());
    for (long long i = 0; i < n; ++i) if (td[i]) { h->has_tdyn = true; break; }
  }
  EPH_CUDA(h, cudaMemcpyAsync(h->flag.p, fl.data(), n * sizeof(short), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->t_dyn.p, td.data(), n * sizeof(unsigned short), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  h->has_walls = false;
  for (long long i = 0; i < n; ++i) if (fl[i] == 2) { h->has_walls = true; break; }
  // the TMA needs 16-byte global strides (nx even); tiny grids gain nothing from tiles
  static const bool tma_off = std::getenv("EPH_B200_NO_TMA") != nullptr;
  h->tma_ok = !tma_off && (nx % 2 == 0) && nx >= 16 && (long long)ny * nz >= 16 &&
              encode_grid_map(&h->map_T[0], h->T[0].p, nx, ny, nz) && encode_grid_map(&h->map_T[1], h->T[1].p, nx, ny, nz) &&
              encode_grid_map(&h->map_K, h->kappa_e.p, nx, ny, nz);
  if (h->tma_ok) cudaFuncSetAttribute(fdm_substep_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
  // constant-coefficient grid?  (what the `NX NY NZ NULL` constructor creates)
  h->uniform = !h->has_tdyn && !h->has_walls;
  for (long long i = 0; i < n && h->uniform; ++i)
    h->uniform = fl[i] == 1 && rho_e[i] == rho_e[0] && C_e[i] == C_e[0] && kappa_e[i] == kappa_e[0] && (S_e ? S_e[i] == S_e[0] : true);
  h->u_kappa = kappa_e[0]; h->u_rho = rho_e[0]; h->u_C = C_e[0]; h->u_S = S_e ? S_e[0] : 0.0;
  h->map_S_base = nullptr;
  if (h->uniform && h->tma_ok) cudaFuncSetAttribute(fdm_uniform_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUniSmemBytes);
  h->grid_set = true;
  h->minmax_valid = false;
  return EPH_B200_OK;
}

int eph_b200_set_grid_tables(eph_b200_handle *h, int n_T, double dT, const double *coeff_C_e_T,
                             const double *coeff_kappa_e_T, const double *E_e_T) {
  if (!h) return EPH_B200_ERR_ARG;
  if (n_T < 4 || !(dT > 0) || !coeff_C_e_T || !coeff_kappa_e_T || !E_e_T) return fail(h, EPH_B200_ERR_ARG, "set_grid_tables: bad table");
  cudaSetDevice(h->cfg.device);
  EPH_CUDA(h, h->C_T_tab.reserve(2 * (size_t)n_T)); EPH_CUDA(h, h->K_T_tab.reserve(2 * (size_t)n_T)); EPH_CUDA(h, h->E_T_tab.reserve(n_T));
  EPH_CUDA(h, cudaMemcpyAsync(h->C_T_tab.p, coeff_C_e_T, 4 * (size_t)n_T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->K_T_tab.p, coeff_kappa_e_T, 4 * (size_t)n_T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->E_T_tab.p, E_e_T, (size_t)n_T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_T = n_T; h->dT_tab = dT;
  return EPH_B200_OK;
}

static double *grid_field(eph_b200_handle *h, int which) {
  switch (which) {
    case 0: return h->T[h->cur].p; case 1: return h->S_e.p; case 2: return h->rho_e.p;
    case 3: return h->C_e.p; case 4: return h->kappa_e.p; case 5: return h->dT_e_ext ? h->dT_e_ext : h->dT_e.p; default: return nullptr;
  }
}

int eph_b200_get_grid(eph_b200_handle *h, int which, double *out) {
  if (!h || !out) return EPH_B200_ERR_ARG;
  if (!h->grid_set) return fail(h, EPH_B200_ERR_ARG, "get_grid: no grid");
  double *p = grid_field(h, which);
  if (!p) return fail(h, EPH_B200_ERR_ARG, "get_grid: bad field id %d", which);
  cudaSetDevice(h->cfg.device);
  join_grid_stream(h);
  EPH_CUDA(h, cudaMemcpyAsync(out, p, h->ncell * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPH_B200_OK;
}

int eph_b200_put_grid(eph_b200_handle *h, int which, const double *in) {
  if (!h || !in) return EPH_B200_ERR_ARG;
  if (!h->grid_set) return fail(h, EPH_B200_ERR_ARG, "put_grid: no grid");
  double *p = grid_field(h, which);
  if (!p) return fail(h, EPH_B200_ERR_ARG, "put_grid: bad field id %d", which);
  cudaSetDevice(h->cfg.device);
  join_grid_stream(h);
  EPH_CUDA(h, cudaMemcpyAsync(p, in, h->ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  if (which >= 1 && which <= 4) h->uniform = false;   // parameters changed cell by cell: general kernel from now on
  h->minmax_valid = false;
  return EPH_B200_OK;
}

int eph_b200_mean_T(eph_b200_handle *h, double *out) {
  if (!h || !out) return EPH_B200_ERR_ARG;
  if (!h->grid_set) return fail(h, EPH_B200_ERR_ARG, "mean_T: no grid");
  cudaSetDevice(h->cfg.device);
  join_grid_stream(h);
  EPH_CUDA(h, cudaMemsetAsync(h->d_scal.p + 1, 0, sizeof(double), h->stream));
  fdm_sum_kernel<<<std::min(blocks_for(h->ncell, 256), 4 * h->sm_count), 256, 0, h->stream>>>(h->ncell, h->T[h->cur].p, h->d_scal.p + 1);
  EPH_LAUNCH_CHECK(h);
  EPH_CUDA(h, cudaMemcpyAsync(h->h_pinned + 1, h->d_scal.p + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  *out = h->h_pinned[1] / (double)h->ncell;
  return EPH_B200_OK;
}

int eph_b200_last_substeps(eph_b200_handle *h, int *out) {
  if (!h || !out) return EPH_B200_ERR_ARG;
  *out = h->last_substeps;
  return EPH_B200_OK;
}

int eph_b200_set_dt(eph_b200_handle *h, double dt, double boltz) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!(dt > 0) || !(boltz > 0)) return fail(h, EPH_B200_ERR_ARG, "set_dt: dt and boltz must be positive");
  h->dt = dt; h->boltz = boltz;
  h->eta = std::sqrt(2.0 * boltz / dt);  // eta_factor, fix_eph.cpp:200, :910
  if (h->coloured) h->zeta = 1.0 - std::exp(-dt / h->tau0);   // zeta_factor, fix_eph_coloured_exp.cpp:686
  h->dt_set = true;
  return EPH_B200_OK;
}

// `fix eph/coloured/exp`: f_EPH and f_RNG of model 4 pass through f <- f_prev (1 - zeta) + zeta f_new with
// zeta = 1 - exp(-dt / tau0) (fix_eph_coloured_exp.cpp:190, :563-569, :619-625); tau0 <= 0 switches the filter off.
int eph_b200_set_colour(eph_b200_handle *h, double tau0) {
  if (!h) return EPH_B200_ERR_ARG;
  if (tau0 > 0.0 && h->cfg.model != EPH_B200_MODEL_PRL) return fail(h, EPH_B200_ERR_MODEL, "set_colour: the memory kernel belongs to eph_model 4");
  cudaSetDevice(h->cfg.device);
  h->coloured = tau0 > 0.0;
  h->tau0 = tau0;
  h->zeta = (h->coloured && h->dt_set) ? 1.0 - std::exp(-h->dt / tau0) : 0.0;
  if (h->coloured && h->atoms_set) {
    const size_t n3 = 3 * std::max<size_t>(h->nlocal, 1);
    if (h->f_dis.cap < n3 || h->f_sto.cap < n3) {
      EPH_CUDA(h, h->f_dis.reserve(n3)); EPH_CUDA(h, h->f_sto.reserve(n3));
      EPH_CUDA(h, cudaMemsetAsync(h->f_dis.p, 0, h->f_dis.cap * sizeof(double), h->stream));
      EPH_CUDA(h, cudaMemsetAsync(h->f_sto.p, 0, h->f_sto.cap * sizeof(double), h->stream));
    }
  }
  return EPH_B200_OK;
}

// the filter's per-atom state [nlocal][3] each: it migrates with the atoms (pack_exchange / copy_arrays,
// fix_eph_coloured_exp.cpp:793-825), so the host fix reads it back after every step and re-registers it after set_atoms
int eph_b200_get_colour_state(eph_b200_handle *h, double *f_dis, double *f_sto, int memspace) {
  if (!h || !f_dis || !f_sto) return EPH_B200_ERR_ARG;
  if (!h->coloured || !h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "get_colour_state: set_colour and set_atoms first");
  cudaSetDevice(h->cfg.device);
  const size_t bytes = 3 * (size_t)h->nlocal * sizeof(double);
  const cudaMemcpyKind kind = memspace == EPH_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (bytes) {
    EPH_CUDA(h, cudaMemcpyAsync(f_dis, h->f_dis.p, bytes, kind, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(f_sto, h->f_sto.p, bytes, kind, h->stream));
    if (memspace != EPH_B200_DEVICE) EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

int eph_b200_set_colour_state(eph_b200_handle *h, const double *f_dis, const double *f_sto, int memspace) {
  if (!h || !f_dis || !f_sto) return EPH_B200_ERR_ARG;
  if (!h->coloured || !h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "set_colour_state: set_colour and set_atoms first");
  cudaSetDevice(h->cfg.device);
  const size_t bytes = 3 * (size_t)h->nlocal * sizeof(double);
  const cudaMemcpyKind kind = memspace == EPH_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (bytes) {
    EPH_CUDA(h, cudaMemcpyAsync(h->f_dis.p, f_dis, bytes, kind, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(h->f_sto.p, f_sto, bytes, kind, h->stream));
    if (memspace != EPH_B200_DEVICE) EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

int eph_b200_set_precision(eph_b200_handle *h, int packed_records) {
  if (!h) return EPH_B200_ERR_ARG;
  if (h->pf_open) return fail(h, EPH_B200_ERR_ARG, "set_precision: not between post_force_begin and post_force_end");
  h->precision = packed_records ? 1 : 0;
  return EPH_B200_OK;
}

int eph_b200_get_precision(eph_b200_handle *h, int *packed_records, double *period, double *position_quantum) {
  if (!h) return EPH_B200_ERR_ARG;
  const bool on = packed_possible(h);
  if (packed_records) *packed_records = on ? 1 : 0;
  if (period) *period = on ? packed_period(h) : 0.0;
  if (position_quantum) *position_quantum = on ? packed_period(h) / kQScale : 0.0;
  return EPH_B200_OK;
}

int eph_b200_set_skin(eph_b200_handle *h, double skin, double inner_skin) {
  if (!h) return EPH_B200_ERR_ARG;
  if (skin < 0.0) return fail(h, EPH_B200_ERR_ARG, "set_skin: negative skin");
  h->skin = skin;
  if (inner_skin >= 0.0) h->inner_skin = inner_skin;   // negative: keep the current setting
  h->inner_wanted = h->inner_skin > 0.0 && h->inner_skin <= skin;
  h->have_inner = false;
  // with a list registered its age is unknown here and its tile storage was sized under the previous setting: the
  // sweeps walk LAMMPS' list until the next set_neighbors applies the new one
  h->inner_enabled = h->neigh_set ? false : h->inner_wanted;
  return EPH_B200_OK;
}

int eph_b200_list_stats(eph_b200_handle *h, long long *inner_builds, long long *fallback_steps) {
  if (!h) return EPH_B200_ERR_ARG;
  if (inner_builds) *inner_builds = h->inner_builds;
  if (fallback_steps) *fallback_steps = h->inner_fallback_steps;
  return EPH_B200_OK;
}

int eph_b200_set_atoms(eph_b200_handle *h, int nlocal, int nghost, const int *type, const int *mask,
                       const int64_t *tag, const int *ghost_owner, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (nlocal < 0 || nghost < 0 || !type || !mask) return fail(h, EPH_B200_ERR_ARG, "set_atoms: bad counts or null arrays");
  cudaSetDevice(h->cfg.device);
  const size_t nt = (size_t)nlocal + nghost;
  const cudaMemcpyKind kind = memspace == EPH_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  EPH_CUDA(h, h->type.reserve(nt)); EPH_CUDA(h, h->mask.reserve(nt)); EPH_CUDA(h, h->tag.reserve(nt));
  EPH_CUDA(h, cudaMemcpyAsync(h->type.p, type, nt * sizeof(int), kind, h->stream));
  EPH_CUDA(h, cudaMemcpyAsync(h->mask.p, mask, nt * sizeof(int), kind, h->stream));
  if (tag) EPH_CUDA(h, cudaMemcpyAsync(h->tag.p, tag, nt * sizeof(long long), kind, h->stream));
  else if (h->cfg.flags & EPH_B200_RANDOM) {
    // without tags the counter-based stream cannot be keyed; injected xi still works
    EPH_CUDA(h, cudaMemsetAsync(h->tag.p, 0, nt * sizeof(long long), h->stream));
  }
  h->has_owner = false;
  if (nghost > 0 && ghost_owner) {
    EPH_CUDA(h, h->owner.reserve(nghost));
    EPH_CUDA(h, cudaMemcpyAsync(h->owner.p, ghost_owner, (size_t)nghost * sizeof(int), kind, h->stream));
    h->has_owner = true;
  }
  EPH_CUDA(h, h->pos4.reserve(nt)); EPH_CUDA(h, h->pv.reserve(2 * nt)); EPH_CUDA(h, h->puz.reserve(3 * nt));
  EPH_CUDA(h, h->rho.reserve(nt)); EPH_CUDA(h, h->W4.reserve(nt));
  EPH_CUDA(h, h->recD.reserve(nt)); EPH_CUDA(h, h->recA.reserve(nt)); EPH_CUDA(h, h->recB.reserve(nt));
  EPH_CUDA(h, h->var.reserve(std::max<size_t>(nlocal, 1)));
  EPH_CUDA(h, h->xref.reserve(nt)); EPH_CUDA(h, h->xref0.reserve(nt)); EPH_CUDA(h, h->icount.reserve(std::max<size_t>(nlocal, 1)));
  h->have_inner = false;
  const size_t nl = std::max<size_t>(nlocal, 1);
  EPH_CUDA(h, h->w.reserve(3 * nl)); EPH_CUDA(h, h->xi.reserve(3 * std::max<size_t>(nt, 1))); EPH_CUDA(h, h->f_eph.reserve(3 * nl));
  EPH_CUDA(h, h->f_rng.reserve(3 * nl)); EPH_CUDA(h, h->array8.reserve(8 * nl));
  // fresh storage is zero, like the reference constructor (fix_eph.cpp:229-238)
  EPH_CUDA(h, cudaMemsetAsync(h->rho.p, 0, nt * sizeof(double), h->stream));
  EPH_CUDA(h, cudaMemsetAsync(h->puz.p, 0, 3 * nt * sizeof(double4), h->stream));
  EPH_CUDA(h, cudaMemsetAsync(h->w.p, 0, 3 * nl * sizeof(double), h->stream));
  EPH_CUDA(h, cudaMemsetAsync(h->xi.p, 0, 3 * std::max<size_t>(nt, 1) * sizeof(double), h->stream));
  EPH_CUDA(h, cudaMemsetAsync(h->f_eph.p, 0, 3 * nl * sizeof(double), h->stream));
  EPH_CUDA(h, cudaMemsetAsync(h->f_rng.p, 0, 3 * nl * sizeof(double), h->stream));
  EPH_CUDA(h, cudaMemsetAsync(h->array8.p, 0, 8 * nl * sizeof(double), h->stream));
  if (h->coloured && (h->f_dis.cap < 3 * nl || h->f_sto.cap < 3 * nl)) {   // new filter storage starts from zero (fix_eph_coloured_exp.cpp:229-230)
    EPH_CUDA(h, h->f_dis.reserve(3 * nl)); EPH_CUDA(h, h->f_sto.reserve(3 * nl));
    EPH_CUDA(h, cudaMemsetAsync(h->f_dis.p, 0, h->f_dis.cap * sizeof(double), h->stream));
    EPH_CUDA(h, cudaMemsetAsync(h->f_sto.p, 0, h->f_sto.cap * sizeof(double), h->stream));
  }
  if (memspace != EPH_B200_DEVICE) EPH_CUDA(h, cudaStreamSynchronize(h->stream));  // host source buffers may be reused by the caller
  h->nlocal = nlocal; h->nghost = nghost;
  h->atoms_set = true;
  h->peratom_valid = false;
  h->split_ready = false;   // the boundary work lists name atoms of the previous registration
  h->ghost_map_set = false; // ... and so does the ghost map
  h->gshift_valid = false;
  h->resident = false; h->res_f_valid = false; h->res_pf_started = false;
  h->neigh_set = false;
  h->forces_valid = false;
  return EPH_B200_OK;
}

int eph_b200_set_neighbors_csr(eph_b200_handle *h, int nlocal, const int64_t *offsets, const int *neigh, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "set_neighbors: call set_atoms first");
  if (nlocal != h->nlocal) return fail(h, EPH_B200_ERR_ARG, "set_neighbors: nlocal %d differs from set_atoms (%d)", nlocal, h->nlocal);
  if (!offsets || (!neigh && nlocal > 0)) return fail(h, EPH_B200_ERR_ARG, "set_neighbors: null list");
  cudaSetDevice(h->cfg.device);
  long long total = 0;
  if (memspace == EPH_B200_DEVICE) {
    // the list length comes down together with the size of the tile storage: one synchronisation for both
    long long *pinned_total = reinterpret_cast<long long *>(h->h_pinned + 6);
    EPH_CUDA(h, cudaMemcpyAsync(pinned_total, offsets + nlocal, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    h->off_ptr = reinterpret_cast<const long long *>(offsets);
    h->neigh_ptr = neigh;
    int rc = prepare_tiles(h, nlocal, 0, pinned_total);
    if (rc) return rc;
    if (*pinned_total < 0) return fail(h, EPH_B200_ERR_ARG, "set_neighbors: negative list length");
    return register_list(h, nlocal, *pinned_total, true, false);
  } else {
    total = offsets[nlocal];
    EPH_CUDA(h, h->off.reserve((size_t)nlocal + 1));
    EPH_CUDA(h, h->neigh.reserve((size_t)std::max<long long>(total, 1)));
    EPH_CUDA(h, cudaMemcpyAsync(h->off.p, offsets, ((size_t)nlocal + 1) * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(h->neigh.p, neigh, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
    h->off_ptr = h->off.p;
    h->neigh_ptr = h->neigh.p;
  }
  if (total < 0) return fail(h, EPH_B200_ERR_ARG, "set_neighbors: negative list length");
  return register_list(h, nlocal, total, false, false);
}

}  // extern "C"

// Common tail of the ways a full list reaches the engine: tile storage of the inner list sized from the rows, pair-weight
// slots, list state.  tiles_ready: the caller has sized the tiles already; inner_filled: eph_b200_build_neighbors has
// written the inner list into them as well.
static int register_list(eph_b200_handle *h, int nlocal, long long total, bool tiles_ready, bool inner_filled) {
  if (!tiles_ready) {
    int rc = prepare_tiles(h, nlocal, total, nullptr);
    if (rc) return rc;
  }
  EPH_CUDA(h, cudaMemsetAsync(h->lstate.p, 0, sizeof(ListState), h->stream));
  h->fresh_neighbors = true;
  h->have_inner = false;
  h->inner_gave_up = false;
  h->flag_pending = false;
  h->inner_prebuilt = inner_filled && h->inner_enabled;
  h->n_entries = total;
  h->neigh_set = true;
  return EPH_B200_OK;
}

// total_pending: the list length is still on its way to the host (pinned); it is there after this function's own
// synchronisation, or after the one done here if the tiles need none
static int prepare_tiles(eph_b200_handle *h, int nlocal, long long total, const long long *total_pending) {
  h->inner_enabled = h->inner_wanted;
  long long slots = total;   // pair-weight slots: CSR rows of LAMMPS' list or, usually larger, the tiles of the inner list
  bool synced = false;
  if (h->inner_enabled && nlocal > 0) {
    const int tile_atoms = 32 / h->lanes;
    const int ntiles = (nlocal + tile_atoms - 1) / tile_atoms;
    EPH_CUDA(h, h->tile_caps.reserve((size_t)ntiles + 1));
    EPH_CUDA(h, h->tile_off.reserve((size_t)ntiles + 1));
    tile_caps_kernel<<<blocks_for(ntiles + 1, 256), 256, 0, h->stream>>>(nlocal, h->off_ptr, tile_atoms, h->lanes, ntiles, h->tile_caps.p);
    EPH_LAUNCH_CHECK(h);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, h->tile_caps.p, h->tile_off.p, ntiles + 1, h->stream);
    EPH_CUDA(h, h->nb_tmp.reserve(scan_bytes));
    EPH_CUDA(h, cub::DeviceScan::ExclusiveSum(h->nb_tmp.p, scan_bytes, h->tile_caps.p, h->tile_off.p, ntiles + 1, h->stream));
    ++h->launches;
    long long *tiled = reinterpret_cast<long long *>(h->h_pinned + 7);
    EPH_CUDA(h, cudaMemcpyAsync(tiled, h->tile_off.p + ntiles, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
    synced = true;
    if (total_pending) slots = *total_pending;
    EPH_CUDA(h, h->ineigh.reserve((size_t)std::max<long long>(*tiled, 1)));
    slots = std::max(slots, *tiled);
  }
  if (total_pending && !synced) {
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
    slots = *total_pending;
  }
  EPH_CUDA(h, h->gpair.reserve((size_t)std::max<long long>(slots, 1)));
  if (h->n_el > 1) EPH_CUDA(h, h->gpair_i.reserve((size_t)std::max<long long>(slots, 1)));
  return EPH_B200_OK;
}

extern "C" {

// Builds the full list on the device from the positions: rows of local atoms over all atoms closer than `cutoff`
// (r_c + skin).  Same pair set as LAMMPS' REQ_FULL list, so the fix need not request (and LAMMPS need not build and
// the host need not upload) that list at all.
int eph_b200_build_neighbors(eph_b200_handle *h, const double *x, double cutoff, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "build_neighbors: call set_atoms first");
  if ((!x && !h->resident) || !(cutoff > 0.0)) return fail(h, EPH_B200_ERR_ARG, "build_neighbors: null positions or bad cut-off");
  if (!x) { x = h->res_x.p; memspace = EPH_B200_DEVICE; }   // the positions the engine keeps (eph_b200_resident_upload)
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal, nt = h->nlocal + h->nghost;
  if (nl == 0) {
    const long long zero = 0;
    return eph_b200_set_neighbors_csr(h, 0, reinterpret_cast<const int64_t *>(&zero), nullptr, EPH_B200_HOST);
  }
  const double *dx = nullptr;
  int rc;
  if ((rc = stage_in(h, h->x, x, 3 * (size_t)nt, memspace, &dx))) return rc;
  KernelTimer kt(h, "build_neighbors");
  // bounding box of locals + ghosts
  EPH_CUDA(h, h->nb_box.reserve(6));
  const double init[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
  double box[6];
  EPH_CUDA(h, cudaMemcpyAsync(h->nb_box.p, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
  bbox_kernel<<<std::min(blocks_for(nt, 256), 8 * h->sm_count), 256, 0, h->stream>>>(nt, dx, h->nb_box.p);
  EPH_LAUNCH_CHECK(h);
  EPH_CUDA(h, cudaMemcpyAsync(box, h->nb_box.p, sizeof box, cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  CellGrid g;
  long long ncell = 1;
  for (int d = 0; d < 3; ++d) {
    const double len = std::max(box[3 + d] - box[d], 1e-9);
    g.nb[d] = std::max(1, (int)std::floor(len / (0.5 * cutoff)));   // cells of at least cutoff/2: +-2 cells cover the cut-off
    g.lo[d] = box[d];
    g.inv[d] = g.nb[d] / (len * (1.0 + 1e-12));
    ncell *= g.nb[d];
  }
  if (ncell > 500000000LL) return fail(h, EPH_B200_ERR_ARG, "build_neighbors: %lld cells (box far larger than the atoms?)", ncell);
  EPH_CUDA(h, h->nb_cell.reserve(nt)); EPH_CUDA(h, h->nb_atom.reserve(nt)); EPH_CUDA(h, h->nb_cell_s.reserve(nt));
  EPH_CUDA(h, h->nb_atom_s.reserve(nt)); EPH_CUDA(h, h->nb_xs.reserve(nt)); EPH_CUDA(h, h->nb_start.reserve(ncell));
  EPH_CUDA(h, h->nb_end.reserve(ncell)); EPH_CUDA(h, h->nb_counts.reserve((size_t)nl + 1)); EPH_CUDA(h, h->off.reserve((size_t)nl + 1));
  cell_id_kernel<<<blocks_for(nt, 256), 256, 0, h->stream>>>(nt, dx, g, h->nb_cell.p, h->nb_atom.p);
  EPH_LAUNCH_CHECK(h);
  int bits = 1;
  while ((1LL << bits) < ncell) ++bits;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->nb_cell.p, h->nb_cell_s.p, h->nb_atom.p, h->nb_atom_s.p, nt, 0, bits, h->stream);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, h->nb_counts.p, h->off.p, nl + 1, h->stream);
  EPH_CUDA(h, h->nb_tmp.reserve(std::max(tmp_bytes, scan_bytes)));
  EPH_CUDA(h, cub::DeviceRadixSort::SortPairs(h->nb_tmp.p, tmp_bytes, h->nb_cell.p, h->nb_cell_s.p, h->nb_atom.p, h->nb_atom_s.p, nt, 0, bits, h->stream));
  ++h->launches;
  EPH_CUDA(h, cudaMemsetAsync(h->nb_start.p, 0, (size_t)ncell * sizeof(int), h->stream));
  EPH_CUDA(h, cudaMemsetAsync(h->nb_end.p, 0, (size_t)ncell * sizeof(int), h->stream));
  cell_ranges_kernel<<<blocks_for(nt, 256), 256, 0, h->stream>>>(nt, h->nb_cell_s.p, h->nb_atom_s.p, dx, h->nb_xs.p, h->nb_start.p, h->nb_end.p);
  EPH_LAUNCH_CHECK(h);
  EPH_CUDA(h, cudaMemsetAsync(h->nb_counts.p + nl, 0, sizeof(long long), h->stream));
  // count and fill passes: one CTA per row of cells, candidates staged through shared memory (EPH_B200_NEIGH_KERNEL=atom
  // selects the per-atom kernels instead)
  static const bool per_atom = std::getenv("EPH_B200_NEIGH_KERNEL") && std::strcmp(std::getenv("EPH_B200_NEIGH_KERNEL"), "atom") == 0;
  const int nrows = g.nb[1] * g.nb[2];
  if (per_atom)
    neighbor_pass_kernel<false><<<blocks_for(nl, 128), 128, 0, h->stream>>>(nl, dx, g, cutoff * cutoff, h->nb_xs.p, h->nb_start.p, h->nb_end.p,
                                                                            h->nb_counts.p, nullptr, nullptr);
  else
    neighbor_tile_kernel<false><<<nrows, 32 * kNeighWarps, 0, h->stream>>>(nl, g, cutoff * cutoff, h->nb_xs.p, h->nb_start.p, h->nb_end.p,
                                                                           h->nb_counts.p, nullptr, nullptr, InnerOut{});
  EPH_LAUNCH_CHECK(h);
  EPH_CUDA(h, cub::DeviceScan::ExclusiveSum(h->nb_tmp.p, scan_bytes, h->nb_counts.p, h->off.p, nl + 1, h->stream));
  ++h->launches;
  long long total = 0;
  EPH_CUDA(h, cudaMemcpyAsync(&total, h->off.p + nl, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  EPH_CUDA(h, h->neigh.reserve((size_t)std::max<long long>(total, 1)));
  if (per_atom)
    neighbor_pass_kernel<true><<<blocks_for(nl, 128), 128, 0, h->stream>>>(nl, dx, g, cutoff * cutoff, h->nb_xs.p, h->nb_start.p, h->nb_end.p,
                                                                           nullptr, h->off.p, h->neigh.p);
  else {
    // With packed records the step after a re-neighbouring runs on the inner list straight away, so the fill pass also
    // writes the inner list (same rows, cut at r_c + inner_skin) into its tiles and no second walk of the new list is needed.
    h->off_ptr = h->off.p;
    h->neigh_ptr = h->neigh.p;
    int rc = prepare_tiles(h, nl, total, nullptr);
    if (rc) return rc;
    InnerOut io{};
    const bool prebuild = h->inner_enabled && packed_possible(h);
    if (prebuild) {
      const double r_in = std::sqrt(h->rc2) + h->inner_skin;
      io.ineigh = h->ineigh.p; io.icount = h->icount.p; io.tile_off = h->tile_off.p; io.mask = h->mask.p;
      io.groupbit = h->cfg.groupbit; io.r_inner_sq = r_in * r_in; io.lanes = h->lanes;
    }
    neighbor_tile_kernel<true><<<nrows, 32 * kNeighWarps, 0, h->stream>>>(nl, g, cutoff * cutoff, h->nb_xs.p, h->nb_start.p, h->nb_end.p,
                                                                          nullptr, h->off.p, h->neigh.p, io);
    EPH_LAUNCH_CHECK(h);
    if (prebuild) {
      const int ntiles = (nl + 32 / h->lanes - 1) / (32 / h->lanes);
      tile_pad_kernel<<<(int)std::min<long long>(blocks_for(ntiles, 8), (long long)h->sm_count * 8), 256, 0, h->stream>>>(
          nl, h->lanes, h->tile_off.p, h->icount.p, h->ineigh.p, h->gpair.p, h->n_el > 1 ? h->gpair_i.p : nullptr);
      EPH_LAUNCH_CHECK(h);
    }
    return register_list(h, nl, total, true, prebuild);
  }
  EPH_LAUNCH_CHECK(h);
  // hand the device-resident CSR to the common path (aliases our own buffers: no copy)
  return eph_b200_set_neighbors_csr(h, nl, reinterpret_cast<const int64_t *>(h->off.p), h->neigh.p, EPH_B200_DEVICE);
}

// read-back of the list currently in use (tests, diagnostics): offsets[nlocal+1]; neigh may be NULL to query the size
int eph_b200_get_neighbors(eph_b200_handle *h, int64_t *offsets, int *neigh, long long *n_entries) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->neigh_set) return fail(h, EPH_B200_ERR_ARG, "get_neighbors: no list");
  cudaSetDevice(h->cfg.device);
  if (n_entries) *n_entries = h->n_entries;
  if (offsets) EPH_CUDA(h, cudaMemcpyAsync(offsets, h->off_ptr, ((size_t)h->nlocal + 1) * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  if (neigh && h->n_entries > 0) EPH_CUDA(h, cudaMemcpyAsync(neigh, h->neigh_ptr, (size_t)h->n_entries * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPH_B200_OK;
}

int eph_b200_set_neighbors_lammps(eph_b200_handle *h, int nlocal, const int *numneigh, int *const *firstneigh) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!numneigh || (!firstneigh && nlocal > 0)) return fail(h, EPH_B200_ERR_ARG, "set_neighbors: null list");
  // LAMMPS keeps the list in paged host memory (int **firstneigh); flatten once per
  // re-neighbouring -- never per step (the legacy port's fix_eph_gpu.cpp:241-271 did).
  std::vector<long long> off((size_t)nlocal + 1, 0);
  for (int i = 0; i < nlocal; ++i) off[i + 1] = off[i] + numneigh[i];
  std::vector<int> flat((size_t)std::max<long long>(off[nlocal], 1));
  for (int i = 0; i < nlocal; ++i) std::memcpy(flat.data() + off[i], firstneigh[i], sizeof(int) * (size_t)numneigh[i]);
  return eph_b200_set_neighbors_csr(h, nlocal, reinterpret_cast<const int64_t *>(off.data()), flat.data(), EPH_B200_HOST);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// sweep launch helpers
// ---------------------------------------------------------------------------
namespace {

int env_int(const char *name, int dflt);

// Grid of a sweep over `items` atoms with `per_cta` atoms per CTA pass: persistent, one resident wave of CTAs
// striding over the atoms.  Measured alternative (EPH_B200_PERSISTENT=0: many short-lived CTAs, so that the kernels of
// the communication and grid streams -- pack/unpack, NCCL, the grid solve -- can take SM slots while a sweep is
// running instead of waiting for its tail): 3 % slower on one GPU and 2.5 % slower on 8 GPUs, where the NCCL kernels
// then spin on SMs the sweep could use (profiles/r1_v10_bench_4M_8gpu*.json).  Kept as a switch, not the default.
// A launch that only stands by for a packed sweep (SweepArgs::only_fallback) is empty in all but the steps in which the
// displacement guard trips: one CTA per SM keeps the empty launch at ~3 us (a full resident wave costs ~6) and the rare
// real one merely runs at reduced occupancy.
template <class K>
int sweep_grid(eph_b200_handle *h, K kernel, int threads, size_t smem, long long items, int per_cta, bool standby = false) {
  static const int forced = env_int("EPH_B200_PERSISTENT", -1);
  const bool persistent = forced >= 0 ? forced != 0 : true;
  const long long passes = std::max<long long>(1, (items + per_cta - 1) / per_cta);
  if (!persistent) return (int)std::min<long long>((passes + 3) / 4, 1 << 22);   // four passes per CTA: short-lived CTAs
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return (int)std::min<long long>(passes, (long long)h->sm_count * (standby ? 1 : per_sm));
}

int env_int(const char *name, int dflt) {
  const char *e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

template <int LANES, int TAB, bool MULTI>
int launch_density(eph_b200_handle *h, const SweepArgs &a, size_t smem, bool build, const char *name) {
  const int threads = 256;
  KernelTimer kt(h, name ? name : (build ? "density_sweep_build" : "density_sweep"));
  if (build) {
    auto k = density_sweep_kernel<LANES, TAB, true, MULTI>;
    if (TAB) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<sweep_grid(h, k, threads, smem, a.n_work, threads / LANES), threads, smem, h->stream>>>(a);
  } else {
    auto k = density_sweep_kernel<LANES, TAB, false, MULTI>;
    if (TAB) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<sweep_grid(h, k, threads, smem, a.n_work, threads / LANES, a.only_fallback != 0), threads, smem, h->stream>>>(a);
  }
  EPH_LAUNCH_CHECK(h);
  return EPH_B200_OK;
}

template <int LANES, bool MULTI>
int launch_force(eph_b200_handle *h, const SweepArgs &a) {
  const int threads = EPH_THREADS_FORCE;
  KernelTimer kt(h, a.only_fallback ? "force_sweep_fallback" : "force_sweep");
  auto k = force_sweep_kernel<LANES, MULTI>;
  k<<<sweep_grid(h, k, threads, 0, a.i_end - a.i_begin, threads / LANES, a.only_fallback != 0), threads, 0, h->stream>>>(a);
  EPH_LAUNCH_CHECK(h);
  return EPH_B200_OK;
}

template <int TAB, bool MULTI>
int launch_density_lanes(eph_b200_handle *h, const SweepArgs &a, size_t smem, int lanes, bool build, const char *name) {
  switch (lanes) {
    case 1: return launch_density<1, TAB, MULTI>(h, a, smem, build, name);
    case 2: return launch_density<2, TAB, MULTI>(h, a, smem, build, name);
    case 8: return launch_density<8, TAB, MULTI>(h, a, smem, build, name);
    case 16: return launch_density<16, TAB, MULTI>(h, a, smem, build, name);
    default: return launch_density<4, TAB, MULTI>(h, a, smem, build, name);
  }
}

// ---- packed records (eph_packed.cuh) ----
PackedArgs packed_args(const eph_b200_handle *h) {
  PackedArgs q{};
  q.D = h->recD.p; q.A = h->recA.p; q.B = h->recB.p; q.pos4 = h->pos4.p; q.var = h->var.p;
  const double quantum = packed_period(h) / kQScale;
  q.quantum_sq = quantum * quantum;
  return q;
}

template <int LANES, bool MULTI, bool FRIC>
int launch_density_packed(eph_b200_handle *h, const SweepArgs &a, const char *name) {
  const int threads = 256;
  KernelTimer kt(h, name ? name : "density_sweep");
  auto k = density_packed_kernel<LANES, MULTI, FRIC>;
  k<<<sweep_grid(h, k, threads, 0, a.n_work, threads / LANES), threads, 0, h->stream>>>(a, packed_args(h));
  EPH_LAUNCH_CHECK(h);
  return EPH_B200_OK;
}
template <int LANES, bool MULTI, bool FRIC, bool RAND>
int launch_force_packed(eph_b200_handle *h, const SweepArgs &a) {
  const int threads = EPH_THREADS_FORCE_PACKED;
  KernelTimer kt(h, "force_sweep");
  auto k = force_packed_kernel<LANES, MULTI, FRIC, RAND>;
  k<<<sweep_grid(h, k, threads, 0, a.i_end - a.i_begin, threads / LANES), threads, 0, h->stream>>>(a, packed_args(h));
  EPH_LAUNCH_CHECK(h);
  return EPH_B200_OK;
}
template <int LANES, bool MULTI>
int launch_packed_flags(eph_b200_handle *h, const SweepArgs &a, int which, const char *name) {
  if (which == 0) return a.do_friction ? launch_density_packed<LANES, MULTI, true>(h, a, name) : launch_density_packed<LANES, MULTI, false>(h, a, name);
  if (a.do_friction && a.do_random) return launch_force_packed<LANES, MULTI, true, true>(h, a);
  if (a.do_friction) return launch_force_packed<LANES, MULTI, true, false>(h, a);
  return launch_force_packed<LANES, MULTI, false, true>(h, a);
}
// which: 0 density, 1 force
int launch_packed(eph_b200_handle *h, const SweepArgs &a, int which, const char *name = nullptr) {
  const bool multi = a.n_elements > 1;
  switch (h->lanes) {
    case 1: return multi ? launch_packed_flags<1, true>(h, a, which, name) : launch_packed_flags<1, false>(h, a, which, name);
    case 2: return multi ? launch_packed_flags<2, true>(h, a, which, name) : launch_packed_flags<2, false>(h, a, which, name);
    case 8: return multi ? launch_packed_flags<8, true>(h, a, which, name) : launch_packed_flags<8, false>(h, a, which, name);
    case 16: return multi ? launch_packed_flags<16, true>(h, a, which, name) : launch_packed_flags<16, false>(h, a, which, name);
    default: return multi ? launch_packed_flags<4, true>(h, a, which, name) : launch_packed_flags<4, false>(h, a, which, name);
  }
}

int launch_sweep_exact(eph_b200_handle *h, const SweepArgs &a, int which, bool build, const char *name);

// which: 0 density pass (optionally rebuilding the inner list), 1 force pass.  Both passes use the same number of
// lanes per atom: it fixes the tile shape of the inner list and of the pair weights.  `packed`: the pass runs on the
// packed records; where the inner list may turn out invalid on the device (density pass, force pass in walk mode 1)
// the fp64 kernel is launched behind it as a fall-back that returns at once in the normal case.
int launch_sweep(eph_b200_handle *h, const SweepArgs &a, int which, bool build = false, const char *name = nullptr,
                 bool packed = false) {
  if (!packed) return launch_sweep_exact(h, a, which, build, name);
  int rc = launch_packed(h, a, which, name);
  if (rc) return rc;
  if (which == 1 && a.walk_mode == 2) return EPH_B200_OK;   // list built in this very step: nothing to fall back from
  SweepArgs b = a;
  b.only_fallback = 1;
  return launch_sweep_exact(h, b, which, false, which == 0 ? "density_sweep_fallback" : "force_sweep_fallback");
}

int launch_sweep_exact(eph_b200_handle *h, const SweepArgs &a, int which, bool build, const char *name) {
  const bool multi = a.n_elements > 1;
  if (which == 1) {
    switch (h->lanes) {
      case 1: return multi ? launch_force<1, true>(h, a) : launch_force<1, false>(h, a);
      case 2: return multi ? launch_force<2, true>(h, a) : launch_force<2, false>(h, a);
      case 8: return multi ? launch_force<8, true>(h, a) : launch_force<8, false>(h, a);
      case 16: return multi ? launch_force<16, true>(h, a) : launch_force<16, false>(h, a);
      default: return multi ? launch_force<4, true>(h, a) : launch_force<4, false>(h, a);
    }
  }
  // rho(r^2) tables of all elements: read with 256-bit loads through L1 (default; measured 3 % faster than the
  // shared-memory copy, whose bank conflicts cost more data-pipe wavefronts than the cached sectors, and it keeps
  // 4 CTAs per SM for multi-element tables), or staged in shared memory when they fit (EPH_B200_TABLE=1)
  const size_t table_bytes = (size_t)a.n_elements * a.n_rho * 2 * sizeof(double2);
  static const int table_mode = env_int("EPH_B200_TABLE", 0);
  const bool smem = table_mode == 1 && table_bytes <= (size_t)h->max_smem_optin - 1024;
  if (smem) {
    if (multi) return launch_density_lanes<1, true>(h, a, table_bytes, h->lanes, build, name);
    return launch_density_lanes<1, false>(h, a, table_bytes, h->lanes, build, name);
  }
  if (multi) return launch_density_lanes<0, true>(h, a, 0, h->lanes, build, name);
  return launch_density_lanes<0, false>(h, a, 0, h->lanes, build, name);
}

SweepArgs sweep_args(eph_b200_handle *h) {
  SweepArgs a{};
  a.nlocal = h->nlocal; a.n_elements = h->n_el; a.n_rho = h->n_rho;
  a.inv_dr_sq = h->inv_dr_sq; a.r_cutoff_sq = h->rc2; a.rho_tab4 = reinterpret_cast<const double4 *>(h->rho_tab.p);
  const double r_in = std::sqrt(h->rc2) + h->inner_skin;
  a.r_inner_sq = r_in * r_in;
  a.offsets = h->off_ptr; a.neigh = h->neigh_ptr;
  a.ineigh = h->ineigh.p; a.tile_off = h->tile_off.p; a.icount = h->icount.p; a.inner_invalid = &h->lstate.p->inner_invalid;
  a.use_inner = 0;
  a.i_begin = 0; a.i_end = h->nlocal;
  static const int spec_v = env_int("EPH_B200_SPEC_V", 1);
  a.spec_v = spec_v;
  a.pv = h->pv.p; a.puz = h->puz.p; a.W4 = h->W4.p; a.rho = h->rho.p;
  a.gpair = h->gpair.p; a.gpair_i = h->gpair_i.p;
  a.f = nullptr; a.f_eph = h->f_eph.p; a.f_rng = h->f_rng.p;
  // the pair sums W (and with them w, u) belong to model PRL; the legacy models leave w_i zero like the reference
  a.do_friction = ((h->cfg.flags & EPH_B200_FRICTION) && h->cfg.model == EPH_B200_MODEL_PRL) ? 1 : 0;
  a.do_random = (h->cfg.flags & EPH_B200_RANDOM) ? 1 : 0;
  return a;
}

}  // namespace

extern "C" {

int eph_b200_post_force_begin(eph_b200_handle *h, const double *x, const double *v, const double *xi_inject,
                              long long ntimestep, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->tables_set) return fail(h, EPH_B200_ERR_ARG, "post_force: set_tables not called");
  if (!h->dt_set) return fail(h, EPH_B200_ERR_ARG, "post_force: set_dt not called");
  if (!h->atoms_set || !h->neigh_set) return fail(h, EPH_B200_ERR_ARG, "post_force: set_atoms / set_neighbors not called");
  if ((h->cfg.flags & EPH_B200_RANDOM) && !h->grid_set) return fail(h, EPH_B200_ERR_ARG, "post_force: random force needs the T_e grid");
  if (!x || !v) return fail(h, EPH_B200_ERR_ARG, "post_force: null x or v");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal, nt = h->nlocal + h->nghost;
  h->pf_open = false;
  if (nl == 0) { h->pf_open = true; return EPH_B200_OK; }
  const double *dx = nullptr, *dv = nullptr, *dxi = nullptr;
  int rc;
  if ((rc = stage_in(h, h->x, x, 3 * (size_t)nt, memspace, &dx))) return rc;
  if ((rc = stage_in(h, h->v, v, 3 * (size_t)nt, memspace, &dv))) return rc;
  if (xi_inject && (h->cfg.flags & EPH_B200_RANDOM)) {
    if ((rc = stage_in(h, h->xi_in, xi_inject, 3 * (size_t)nl, memspace, &dxi))) return rc;
  }

  // ---- two-level list policy (host side; correctness is guarded by the device flag, not by this) ----
  bool build = false;
  if (h->inner_enabled) {
    if (h->flag_pending && cudaEventQuery(h->flag_event) == cudaSuccess) {
      h->flag_pending = false;
      if (*h->h_flag != 0u) {
        if (h->rebuilt_last_step) h->inner_gave_up = true;   // the device refused the rebuilt list: stop trying
        else build = !h->inner_gave_up;
        ++h->inner_fallback_steps;
      }
    }
    if (h->fresh_neighbors) build = true;
  }
  const bool track = h->inner_enabled && h->have_inner;
  // the displacement since LAMMPS built its list only matters when the inner list is rebuilt from that (aged) list:
  // the new inner list is complete iff r_c + inner_skin + 2 D(now) <= r_c + skin (prep_coupling decides on the device)
  const bool track0 = track && build && !h->fresh_neighbors;
  if (track0)
    EPH_CUDA(h, cudaMemsetAsync(&h->lstate.p->disp0_sq_bits, 0, sizeof(unsigned long long), h->stream));
  // packed records serve every step that walks the inner list.  A step that has to (re)build it does that first, with
  // a kernel of its own, and then runs the packed passes like any other step; with fp64 records the density pass
  // builds the list while it walks LAMMPS' list.
  const bool build_first = build && packed_possible(h);
  const bool use_inner = h->inner_enabled && (build_first || (h->have_inner && !build));
  const bool packed = packed_possible(h) && use_inner;
  h->step_packed = packed;
  {
    KernelTimer kt(h, "pack_atoms");
    pack_atoms_kernel<<<blocks_for(nt, 256), 256, 0, h->stream>>>(nt, dx, dv, h->type.p, h->mask.p, h->d_type_map.p,
                                                                   h->cfg.groupbit, h->pos4.p, packed ? nullptr : h->pv.p, track ? 1 : 0,
                                                                   track0 ? 1 : 0, h->xref.p, h->xref0.p, h->lstate.p,
                                                                   packed ? h->recD.p : nullptr, 1.0 / packed_period(h), h->d_status.p);
  }
  EPH_LAUNCH_CHECK(h);
  SweepArgs a = sweep_args(h);
  if (build_first && h->inner_prebuilt && h->fresh_neighbors) {
    // eph_b200_build_neighbors wrote the inner list together with the full one
    EPH_CUDA(h, cudaMemsetAsync(&h->lstate.p->inner_invalid, 0, sizeof(unsigned), h->stream));
  } else if (build_first) {
    KernelTimer kt(h, "inner_list_build");
    const int ntiles = (nl + 32 / h->lanes - 1) / (32 / h->lanes);
    const int grid = (int)std::min<long long>(blocks_for(ntiles, 8), (long long)h->sm_count * 8);
    const bool multi = h->n_el > 1;
    switch (h->lanes) {
#define EPH_BUILD_CASE(L) case L: if (multi) inner_build_kernel<L, true><<<grid, 256, 0, h->stream>>>(a, h->pos4.p); \
                                  else inner_build_kernel<L, false><<<grid, 256, 0, h->stream>>>(a, h->pos4.p); break;
      EPH_BUILD_CASE(1) EPH_BUILD_CASE(2) EPH_BUILD_CASE(8) EPH_BUILD_CASE(16)
      default: EPH_BUILD_CASE(4)
#undef EPH_BUILD_CASE
    }
    EPH_LAUNCH_CHECK(h);
    // the list is new: whatever the displacement guard said about the old one no longer applies
    EPH_CUDA(h, cudaMemsetAsync(&h->lstate.p->inner_invalid, 0, sizeof(unsigned), h->stream));
  } else if (packed) {   // fp64 records only if the guard just tripped (the kernel returns at once otherwise)
    pv_fill_kernel<<<std::min(blocks_for(nt, 256), 8 * h->sm_count), 256, 0, h->stream>>>(nt, dv, h->pos4.p, h->pv.p, h->lstate.p);
    EPH_LAUNCH_CHECK(h);
  }
  const bool build_in_pass = build && !build_first;
  a.use_inner = use_inner ? 1 : 0;
  const int tile_atoms = 32 / h->lanes;
  a.work = nullptr; a.n_work = nl; a.n_boundary = 0; a.done_counter = nullptr;
  h->boundary_by_counter = false;
  if (h->comm_stream && h->split_ready && h->n_first > 0 && stream_wait_value() != nullptr) {
    // ONE launch over the reordered tiles, boundary tiles first; each finished boundary tile bumps a counter the
    // communication stream waits on (stream memory operation), so the exchange starts while this launch is still
    // sweeping the interior tiles
    a.work = h->work_all.p; a.n_work = (h->n_first + h->n_rest) * tile_atoms;
    a.n_boundary = h->n_first * tile_atoms; a.done_counter = h->done_counter.p;
    h->boundary_target += (unsigned)h->n_first;
    h->boundary_by_counter = true;
    if ((rc = launch_sweep(h, a, 0, build_in_pass, nullptr, packed))) return rc;
  } else if (h->comm_stream && h->split_ready && h->n_first > 0) {
    // no stream memory operations: two launches with an event between them
    a.work = h->work_all.p; a.n_work = h->n_first * tile_atoms;
    if ((rc = launch_sweep(h, a, 0, build_in_pass, "density_sweep_boundary", packed))) return rc;
    EPH_CUDA(h, cudaEventRecord(h->ev_boundary, h->stream));
    if (h->n_rest > 0) {
      a.work = h->work_all.p + h->n_first; a.n_work = h->n_rest * tile_atoms;
      if ((rc = launch_sweep(h, a, 0, build_in_pass, nullptr, packed))) return rc;
    }
  } else {
    if ((rc = launch_sweep(h, a, 0, build_in_pass, nullptr, packed))) return rc;
    if (h->comm_stream) EPH_CUDA(h, cudaEventRecord(h->ev_boundary, h->stream));
  }
  h->boundary_recorded = h->comm_stream != nullptr;
  if (build) {
    EPH_CUDA(h, cudaMemcpyAsync(h->xref.p, h->pos4.p, (size_t)nt * sizeof(double4), cudaMemcpyDeviceToDevice, h->stream));
    if (h->fresh_neighbors)
      EPH_CUDA(h, cudaMemcpyAsync(h->xref0.p, h->pos4.p, (size_t)nt * sizeof(double4), cudaMemcpyDeviceToDevice, h->stream));
    h->have_inner = true;
    ++h->inner_builds;
  }
  h->rebuilt_last_step = build && !h->fresh_neighbors;
  h->fresh_neighbors = false;
  h->inner_prebuilt = false;
  if (dxi) EPH_CUDA(h, cudaMemcpyAsync(h->xi.p, dxi, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->pf_build = build;
  h->pf_xi = dxi;
  h->pf_step = ntimestep;
  h->pf_open = true;
  return EPH_B200_OK;
}

int eph_b200_post_force_end(eph_b200_handle *h, double *f, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->pf_open) return fail(h, EPH_B200_ERR_ARG, "post_force_end without post_force_begin");
  h->pf_open = false;
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal, nt = h->nlocal + h->nghost;
  if (nl == 0) { h->forces_valid = true; return EPH_B200_OK; }   // a rank without atoms still takes part in end_of_step
  if (!f) return fail(h, EPH_B200_ERR_ARG, "post_force: null f");
  int rc;
  const bool add_fric = (h->cfg.flags & EPH_B200_FRICTION) && !(h->cfg.flags & EPH_B200_NOFRICTION);
  const bool add_rand = (h->cfg.flags & EPH_B200_RANDOM) && !(h->cfg.flags & EPH_B200_NORANDOM);
  double *df = f;
  if (memspace != EPH_B200_DEVICE) {
    EPH_CUDA(h, h->f.reserve(3 * (size_t)nl));
    if (h->f_prefetched) EPH_CUDA(h, cudaStreamWaitEvent(h->stream, h->f_event, 0));   // uploaded behind the density pass
    else if (add_fric || add_rand) EPH_CUDA(h, cudaMemcpyAsync(h->f.p, f, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    df = h->f.p;
  }
  h->f_prefetched = false;
  const bool build = h->pf_build;
  join_grid_stream(h);   // the force pass reads T_e
  if (h->unpack_pending) {
    EPH_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_unpacked, 0));
    h->unpack_pending = false;
  }
  h->boundary_recorded = false;

  PrepArgs p{};
  p.nlocal = nl; p.ntotal = nt; p.owner = h->has_owner ? h->owner.p : nullptr; p.tag = h->tag.p;
  p.xi_inject = h->pf_xi; p.alpha_tab = h->alpha_tab.p; p.n_beta = h->n_beta; p.inv_drho = h->inv_drho; p.rho_cutoff = h->rho_cut;
  p.seed = h->cfg.seed; p.step = (unsigned long long)h->pf_step; p.do_random = (h->cfg.flags & EPH_B200_RANDOM) ? 1 : 0;
  p.rho = h->rho.p; p.W4 = h->W4.p; p.pos4 = h->pos4.p; p.puz = h->puz.p;
  p.T_e = h->grid_set ? h->T[h->cur].p : nullptr; p.grid = grid_geom(h); p.eta_factor = h->eta;
  p.w = h->w.p; p.xi = h->xi.p; p.status = h->d_status.p;
  p.built_inner = build ? 1 : 0; p.skin = h->skin >= 0.0 ? h->skin : h->inner_skin; p.inner_skin = h->inner_skin;
  p.list_state = h->lstate.p;
  // which list the force pass walks: the one whose slots this step's density pass filled with pair weights
  const int walk_mode = build ? 2 : ((h->inner_enabled && h->have_inner) ? 1 : 0);
  // ... and on which records: packed whenever that is the inner list (its pairs are within reach of the packed
  // positions by construction); the fp64 record is then only written if the guard tripped this step
  const bool force_packed = packed_possible(h) && walk_mode != 0;
  p.recA = force_packed ? h->recA.p : nullptr; p.recB = h->recB.p; p.var = h->var.p; p.inv_period = 1.0 / packed_period(h);
  p.puz_mode = !force_packed ? 2 : (walk_mode == 1 ? 1 : 0);
  {
    KernelTimer kt(h, "prep_coupling");
    prep_coupling_kernel<<<blocks_for(nt, 256), 256, 0, h->stream>>>(p);
  }
  EPH_LAUNCH_CHECK(h);
  if (h->inner_enabled && !h->flag_pending) {
    EPH_CUDA(h, cudaMemcpyAsync(h->h_flag, &h->lstate.p->inner_invalid, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaEventRecord(h->flag_event, h->stream));
    h->flag_pending = true;
  }

  SweepArgs a = sweep_args(h);
  if (h->cfg.model == EPH_B200_MODEL_TTM || h->cfg.model == EPH_B200_MODEL_PRB) {
    const bool fric = (h->cfg.flags & EPH_B200_FRICTION) != 0, rnd = (h->cfg.flags & EPH_B200_RANDOM) != 0;
    if (fric || rnd) {
      if (h->cfg.model == EPH_B200_MODEL_PRB && fric) {
        if (!h->rho_r_set) return fail(h, EPH_B200_ERR_ARG, "post_force: eph_model 2 needs eph_b200_set_rho_r_table");
        a.walk_mode = build ? 2 : ((h->inner_enabled && h->have_inner) ? 1 : 0);
        KernelTimer kt(h, "prb_sweep");
        switch (h->lanes) {
          case 1: prb_sweep_kernel<1><<<h->sm_count * 8, 256, 0, h->stream>>>(a, h->rho_r_tab.p, h->inv_dr); break;
          case 2: prb_sweep_kernel<2><<<h->sm_count * 8, 256, 0, h->stream>>>(a, h->rho_r_tab.p, h->inv_dr); break;
          case 8: prb_sweep_kernel<8><<<h->sm_count * 8, 256, 0, h->stream>>>(a, h->rho_r_tab.p, h->inv_dr); break;
          case 16: prb_sweep_kernel<16><<<h->sm_count * 8, 256, 0, h->stream>>>(a, h->rho_r_tab.p, h->inv_dr); break;
          default: prb_sweep_kernel<4><<<h->sm_count * 8, 256, 0, h->stream>>>(a, h->rho_r_tab.p, h->inv_dr); break;
        }
        EPH_LAUNCH_CHECK(h);
      }
      LegacyArgs q{};
      q.nlocal = nl; q.model = h->cfg.model; q.pv = h->pv.p; q.rho = h->rho.p; q.S4 = h->W4.p; q.xi = h->xi.p;
      q.alpha_tab = h->alpha_tab.p; q.beta_tab = h->beta_tab.p; q.n_beta = h->n_beta; q.inv_drho = h->inv_drho;
      q.rho_cutoff = h->rho_cut; q.T_e = h->grid_set ? h->T[h->cur].p : nullptr; q.grid = grid_geom(h);
      q.eta_factor = h->eta; q.do_friction = fric ? 1 : 0; q.do_random = rnd ? 1 : 0;
      q.add_friction = add_fric ? 1 : 0; q.add_random = add_rand ? 1 : 0;
      q.f = (add_fric || add_rand) ? df : nullptr; q.f_eph = h->f_eph.p; q.f_rng = h->f_rng.p;
      {
        KernelTimer kt(h, "legacy_force");
        legacy_force_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(q);
      }
      EPH_LAUNCH_CHECK(h);
      if ((add_fric || add_rand) && memspace != EPH_B200_DEVICE) {
        EPH_CUDA(h, cudaMemcpyAsync(f, h->f.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        EPH_CUDA(h, cudaStreamSynchronize(h->stream));
      }
    }
  } else if (h->cfg.model == EPH_B200_MODEL_PRL && (a.do_friction || a.do_random)) {
    a.f = (add_fric || add_rand) ? df : nullptr;
    a.walk_mode = walk_mode;
    a.add_friction = add_fric ? 1 : 0;
    a.add_random = add_rand ? 1 : 0;
    if (h->coloured) {
      // fix eph/coloured/exp: the force pass only forms f_EPH / f_RNG; the memory kernel filters them and the filtered
      // forces are what goes into f (fix_eph_coloured_exp.cpp:563-569, :619-625, :664-678)
      a.f = nullptr;
      a.i_begin = 0; a.i_end = nl;
      if ((rc = launch_sweep(h, a, 1, false, nullptr, force_packed))) return rc;
      ColourArgs c{};
      c.nlocal = nl; c.pos4 = h->pos4.p; c.rho = h->rho.p; c.zeta = h->zeta;
      c.do_friction = a.do_friction; c.do_random = a.do_random;
      c.add_friction = add_fric ? 1 : 0; c.add_random = add_rand ? 1 : 0;
      c.f_eph = h->f_eph.p; c.f_rng = h->f_rng.p; c.f_dis = h->f_dis.p; c.f_sto = h->f_sto.p;
      c.f = (add_fric || add_rand) ? df : nullptr;
      {
        KernelTimer kt(h, "colour_filter");
        colour_filter_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(c);
      }
      EPH_LAUNCH_CHECK(h);
      if ((add_fric || add_rand) && memspace != EPH_B200_DEVICE) {
        EPH_CUDA(h, cudaMemcpyAsync(f, h->f.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        EPH_CUDA(h, cudaStreamSynchronize(h->stream));
      }
    } else if ((add_fric || add_rand) && memspace != EPH_B200_DEVICE) {
      // host memspace: the force pass runs in a few launches over consecutive atom ranges and every finished range of
      // f goes back to the host on the copy stream while the next range is still being swept
      static const int chunks_env = env_int("EPH_B200_F_CHUNKS", 4);
      const int chunks = std::max(1, std::min(chunks_env, nl / 32768 + 1));
      const int per = ((nl + chunks - 1) / chunks + 255) / 256 * 256;   // multiple of the tile and CTA pass sizes
      for (int i0 = 0; i0 < nl; i0 += per) {
        const int i1 = std::min(nl, i0 + per);
        a.i_begin = i0; a.i_end = i1;
        if ((rc = launch_sweep(h, a, 1, false, nullptr, force_packed))) return rc;
        EPH_CUDA(h, cudaEventRecord(h->f_event, h->stream));
        EPH_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->f_event, 0));
        EPH_CUDA(h, cudaMemcpyAsync(f + 3 * (size_t)i0, h->f.p + 3 * (size_t)i0, 3 * (size_t)(i1 - i0) * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->copy_stream));
      }
      EPH_CUDA(h, cudaStreamSynchronize(h->copy_stream));
      EPH_CUDA(h, cudaStreamSynchronize(h->stream));
    } else {
      a.i_begin = 0; a.i_end = nl;
      if ((rc = launch_sweep(h, a, 1, false, nullptr, force_packed))) return rc;
    }
  }
  h->forces_valid = true;
  return EPH_B200_OK;
}

int eph_b200_post_force(eph_b200_handle *h, const double *x, const double *v, double *f, const double *xi_inject,
                        long long ntimestep, int memspace) {
  if (h && !f && h->nlocal > 0) return fail(h, EPH_B200_ERR_ARG, "post_force: null x, v or f");
  int rc = eph_b200_post_force_begin(h, x, v, xi_inject, ntimestep, memspace);
  if (rc) return rc;
  if (memspace != EPH_B200_DEVICE && h->nlocal > 0 && h->pf_open) {
    // f is only needed by the force pass: send it up on the copy stream while the density pass runs
    const bool add = ((h->cfg.flags & EPH_B200_FRICTION) && !(h->cfg.flags & EPH_B200_NOFRICTION)) ||
                     ((h->cfg.flags & EPH_B200_RANDOM) && !(h->cfg.flags & EPH_B200_NORANDOM));
    if (add) {
      EPH_CUDA(h, h->f.reserve(3 * (size_t)h->nlocal));
      EPH_CUDA(h, cudaMemcpyAsync(h->f.p, f, 3 * (size_t)h->nlocal * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
      EPH_CUDA(h, cudaEventRecord(h->f_event, h->copy_stream));
      h->f_prefetched = true;
    }
  }
  // several ranks with the engine's own transport: the one ghost exchange of the step happens here
  if (h->comm && h->comm_size > 1 && (rc = eph_b200_exchange_ghosts(h))) return rc;
  return eph_b200_post_force_end(h, f, memspace);
}

// The stream a step's ghost exchange runs on: the main stream, or the communication stream made to wait for the boundary
// tiles of the density pass (a counter the launch bumps, or an event behind a launch of its own).
static int exchange_stream(eph_b200_handle *h, cudaStream_t *st) {
  *st = h->stream;
  if (!h->comm_stream) return EPH_B200_OK;
  *st = h->comm_stream;
  if (h->boundary_recorded && h->boundary_by_counter) {
    if (stream_wait_value()(*st, (CUdeviceptr)h->done_counter.p, h->boundary_target, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
      return fail(h, EPH_B200_ERR_CUDA, "ghost exchange: cuStreamWaitValue32 failed");
  } else if (h->boundary_recorded) {
    EPH_CUDA(h, cudaStreamWaitEvent(*st, h->ev_boundary, 0));
  }
  return EPH_B200_OK;
}

// Ghost payload of the one exchange a step needs: {rho, Wx, Wy, Wz} per atom, device buffers.
int eph_b200_pack_ghost_payload(eph_b200_handle *h, int n, const int *send_index_dev, double *buf_dev) {
  if (!h) return EPH_B200_ERR_ARG;
  if (n < 0 || (n > 0 && (!send_index_dev || !buf_dev))) return fail(h, EPH_B200_ERR_ARG, "pack_ghost_payload: bad arguments");
  if (n == 0) return EPH_B200_OK;
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = h->stream;
  int rc = exchange_stream(h, &st);
  if (rc) return rc;
  {
    KernelTimer kt(h, "pack_payload", st);
    pack_payload_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, send_index_dev, h->rho.p, h->W4.p, reinterpret_cast<double4 *>(buf_dev));
  }
  EPH_LAUNCH_CHECK(h);
  return EPH_B200_OK;
}

int eph_b200_unpack_ghost_payload(eph_b200_handle *h, int n, const int *recv_index_dev, const double *buf_dev) {
  if (!h) return EPH_B200_ERR_ARG;
  if (n < 0 || (n > 0 && (!recv_index_dev || !buf_dev))) return fail(h, EPH_B200_ERR_ARG, "unpack_ghost_payload: bad arguments");
  if (n == 0) return EPH_B200_OK;
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = h->comm_stream ? h->comm_stream : h->stream;
  {
    KernelTimer kt(h, "unpack_payload", st);
    unpack_payload_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, recv_index_dev, h->rho.p, h->W4.p,
                                                               reinterpret_cast<const double4 *>(buf_dev), h->nlocal + h->nghost);
  }
  EPH_LAUNCH_CHECK(h);
  if (h->comm_stream) {   // post_force_end waits for the ghosts' {rho, W}
    EPH_CUDA(h, cudaEventRecord(h->ev_unpacked, st));
    h->unpack_pending = true;
  }
  return EPH_B200_OK;
}

}  // extern "C"

namespace {

// EPH_FDM::solve (eph_fdm.h:267-400) in two parts so that a multi-rank caller can put a halo exchange between the
// sub-steps.  grid_plan: refresh of temperature-dependent cells and the sub-step count, decided on the host with the
// reference's exact double arithmetic and truncating cast from three device-reduced scalars (cached while the
// parameters are constant).  grid_substep: one explicit sub-step over the planes [z_begin, z_end).
int grid_plan(eph_b200_handle *h, cudaStream_t st) {
  const long long n = h->ncell;
  if (h->has_tdyn) {
    if (h->n_T < 4) return fail(h, EPH_B200_ERR_ARG, "solve: grid has temperature-dependent cells but no parameter tables");
    fdm_refresh_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, h->T[h->cur].p, h->t_dyn.p, h->C_T_tab.p, h->K_T_tab.p,
                                                                   1. / h->dT_tab, h->C_e.p, h->kappa_e.p);
    EPH_LAUNCH_CHECK(h);
    h->minmax_valid = false;
  }
  if (!h->minmax_valid) {
    // seed with cell 0 (eph_fdm.h:290-292)
    EPH_CUDA(h, cudaMemcpyAsync(h->d_mm.p + 0, h->C_e.p, sizeof(double), cudaMemcpyDeviceToDevice, st));
    EPH_CUDA(h, cudaMemcpyAsync(h->d_mm.p + 1, h->rho_e.p, sizeof(double), cudaMemcpyDeviceToDevice, st));
    EPH_CUDA(h, cudaMemcpyAsync(h->d_mm.p + 2, h->kappa_e.p, sizeof(double), cudaMemcpyDeviceToDevice, st));
    fdm_minmax_kernel<<<std::min(blocks_for(n, 256), 4 * h->sm_count), 256, 0, st>>>(
        n, h->C_e.p, h->rho_e.p, h->kappa_e.p, h->flag.p, h->d_mm.p);
    EPH_LAUNCH_CHECK(h);
    EPH_CUDA(h, cudaMemcpyAsync(h->h_pinned + 2, h->d_mm.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    EPH_CUDA(h, cudaStreamSynchronize(st));
    h->c_min = h->h_pinned[2]; h->rho_min = h->h_pinned[3]; h->kappa_max = h->h_pinned[4];
    h->minmax_valid = true;
  }
  const double dx = h->gdx, dy = h->gdy, dz = h->gdz, dt = h->dt;
  double inner_dt = dt / h->steps;
  const double dtdxdydz = inner_dt * (1.0 / dx / dx + 1.0 / dy / dy + 1.0 / dz / dz);
  const double r = dtdxdydz / h->c_min / h->rho_min * h->kappa_max;
  unsigned int new_steps = (unsigned int)h->steps;
  if (r > 0.4) {  // eph_fdm.h:308-313, truncating cast kept
    inner_dt = 0.4 * inner_dt / r;
    new_steps = std::max(static_cast<unsigned int>(dt / inner_dt), 1u);
    inner_dt = dt / new_steps;
  }
  h->last_substeps = (int)new_steps;
  h->plan_inner_dt = inner_dt;
  return EPH_B200_OK;
}

int grid_substep(eph_b200_handle *h, cudaStream_t st, int z_begin, int z_end, bool clear_source) {
  const long long n = h->ncell;
  const double dx = h->gdx, dy = h->gdy, dz = h->gdz;
  GridArgs g{};
  g.nx = h->nx; g.ny = h->ny; g.nz = h->nz; g.ncell = n;
  g.dT_e = h->dT_e_ext ? h->dT_e_ext : h->dT_e.p; g.S_e = h->S_e.p; g.rho_e = h->rho_e.p; g.C_e = h->C_e.p; g.kappa_e = h->kappa_e.p;
  g.flag = h->flag.p; g.t_dyn = h->t_dyn.p;
  g.E_e_T = h->has_tdyn ? h->E_T_tab.p : nullptr; g.n_T = h->n_T; g.dT = h->dT_tab;
  g.inv_dx2 = 1.0 / (dx * dx); g.inv_dy2 = 1.0 / (dy * dy); g.inv_dz2 = 1.0 / (dz * dz);
  g.inner_dt = h->plan_inner_dt; g.status = h->d_status.p;
  g.z_begin = z_begin; g.z_end = z_end;
  const int nzs = z_end - z_begin;
  dim3 block(32, 4, 2);
  dim3 grid((h->nx + block.x - 1) / block.x, (h->ny + block.y - 1) / block.y, (nzs + block.z - 1) / block.z);
  dim3 tgrid((h->nx + kTX - 1) / kTX, (h->ny + kTY - 1) / kTY, (nzs + kTZ - 1) / kTZ);
  g.T_in = h->T[h->cur].p; g.T_out = h->T[1 - h->cur].p;
  g.clear_source = clear_source ? 1 : 0;
  {
    KernelTimer kt(h, "fdm_substep", st);
    if (h->tma_ok && h->uniform) {
      if (h->map_S_base != g.dT_e) {   // the source array may be caller-owned (bind_grid_source)
        if (!encode_grid_map(&h->map_S, g.dT_e, h->nx, h->ny, h->nz, kTX, kTY, kTZ)) return fail(h, EPH_B200_ERR_CUDA, "solve: cannot encode the source tensor map");
        h->map_S_base = g.dT_e;
      }
      GridUniformArgs u;
      u.nx = h->nx; u.ny = h->ny; u.nz = h->nz; u.T_in = g.T_in; u.T_out = g.T_out; u.dT_e = g.dT_e;
      u.kappa = h->u_kappa; u.S = h->u_S; u.rho = h->u_rho; u.C = h->u_C;
      u.inv_rho_C = 1.0 / (h->u_rho * h->u_C);
      u.inv_dx2 = g.inv_dx2; u.inv_dy2 = g.inv_dy2; u.inv_dz2 = g.inv_dz2; u.inner_dt = g.inner_dt;
      u.clear_source = g.clear_source; u.status = g.status;
      u.z_begin = z_begin; u.z_end = z_end;
      fdm_uniform_tma_kernel<<<tgrid, 256, kUniSmemBytes, st>>>(h->map_T[h->cur], h->map_S, u);
    } else if (h->tma_ok) {
      GridTmaArgs ta;
      ta.g = g;
      ta.has_walls = h->has_walls ? 1 : 0;
      fdm_substep_tma_kernel<<<tgrid, 256, kTmaSmemBytes, st>>>(h->map_T[h->cur], h->map_K, ta);
    } else {
      fdm_substep_kernel<<<grid, block, 0, st>>>(g);
    }
  }
  EPH_LAUNCH_CHECK(h);
  h->cur = 1 - h->cur;
  return EPH_B200_OK;
}

int grid_solve(eph_b200_handle *h, cudaStream_t st) {
  int rc = grid_plan(h, st);
  if (rc) return rc;
  const int new_steps = h->last_substeps;
  for (int s = 0; s < new_steps; ++s)
    if ((rc = grid_substep(h, st, 0, h->nz, s + 1 == new_steps))) return rc;
  return EPH_B200_OK;
}

}  // namespace

extern "C" {

int eph_b200_end_of_step_begin(eph_b200_handle *h, const double *x, const double *v, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->forces_valid && h->nlocal > 0) return fail(h, EPH_B200_ERR_ARG, "end_of_step: post_force has not run");
  if (!h->grid_set) return fail(h, EPH_B200_ERR_ARG, "end_of_step: set_grid not called");
  if (!v && h->nlocal > 0) return fail(h, EPH_B200_ERR_ARG, "end_of_step: null v");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  const double *dx = nullptr, *dv = nullptr;
  int rc;
  // only the local part is read here; x == NULL: positions are those of the last post_force (Verlet does not move
  // atoms between post_force and end_of_step), so nothing is uploaded for them
  if (x && nl > 0 && (rc = stage_in(h, h->x, x, 3 * (size_t)nl, memspace, &dx))) return rc;
  if (nl > 0 && (rc = stage_in(h, h->v, v, 3 * (size_t)nl, memspace, &dv))) return rc;
  join_grid_stream(h);   // the previous solve clears dT_e
  EPH_CUDA(h, cudaMemsetAsync(h->d_scal.p, 0, sizeof(double), h->stream));
  if (nl > 0) {
    DepositArgs d{};
    d.nlocal = nl; d.x = dx; d.v = dv; d.pos4 = h->pos4.p; d.f_eph = h->f_eph.p; d.f_rng = h->f_rng.p;
    d.dt = h->dt; d.dVdt = h->dV * h->dt;
    d.do_friction = (h->cfg.flags & EPH_B200_FRICTION) ? 1 : 0;
    d.do_random = (h->cfg.flags & EPH_B200_RANDOM) ? 1 : 0;
    d.grid = grid_geom(h); d.dT_e = h->dT_e_ext ? h->dT_e_ext : h->dT_e.p; d.E_sum = h->d_scal.p;
    {
      KernelTimer kt(h, "deposit");
      deposit_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(d);
    }
    EPH_LAUNCH_CHECK(h);
  }
  if (h->grid_stream) {   // whatever runs next on the grid stream (the caller's all-reduce, the solve) comes after the deposit
    EPH_CUDA(h, cudaEventRecord(h->ev_deposit, h->stream));
    EPH_CUDA(h, cudaStreamWaitEvent(h->grid_stream, h->ev_deposit, 0));
  }
  h->eos_open = true;
  h->peratom_valid = true;
  return EPH_B200_OK;
}

static int end_of_step_finish(eph_b200_handle *h, double *E_local, bool solve) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->eos_open) return fail(h, EPH_B200_ERR_ARG, "end_of_step_end without end_of_step_begin");
  h->eos_open = false;
  h->plan_open = false;
  cudaSetDevice(h->cfg.device);
  int rc;
  cudaStream_t st = h->grid_stream ? h->grid_stream : h->stream;
  if (solve && (h->cfg.flags & EPH_B200_FDM)) {
    if ((rc = grid_solve(h, st))) return rc;
  }
  if (h->grid_stream) {
    EPH_CUDA(h, cudaEventRecord(h->ev_solved, h->grid_stream));
    h->solve_pending = true;
  }
  if (E_local) {
    unsigned *status = reinterpret_cast<unsigned *>(h->h_pinned + 5);
    EPH_CUDA(h, cudaMemcpyAsync(h->h_pinned, h->d_scal.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (h->p2p_ok) EPH_CUDA(h, cudaMemcpyAsync(status, h->d_status.p, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
    *E_local = h->h_pinned[0];
    // the step's one host synchronisation is where a peer that never delivered its ghost rows becomes an error
    if (h->p2p_ok && (*status & kStatusP2PTimeout))
      return fail(h, EPH_B200_ERR_COMM, "a peer-memory ghost exchange timed out: a peer rank did not deliver its rows (results of this step are invalid)");
  }
  return EPH_B200_OK;
}

int eph_b200_end_of_step_end(eph_b200_handle *h, double *E_local) { return end_of_step_finish(h, E_local, true); }

// ---- sharded grid solve: the caller drives the sub-steps of one solve and exchanges halo planes between them ----
int eph_b200_grid_plan_substeps(eph_b200_handle *h, int *n_substeps) {
  if (!h || !n_substeps) return EPH_B200_ERR_ARG;
  if (!h->eos_open) return fail(h, EPH_B200_ERR_ARG, "grid_plan_substeps: only valid between end_of_step_begin and end_of_step_end_external");
  *n_substeps = 0;
  h->plan_open = false;
  if (!(h->cfg.flags & EPH_B200_FDM)) return EPH_B200_OK;   // the reference does not solve without flag 4 (fix_eph.cpp:392-394)
  cudaSetDevice(h->cfg.device);
  int rc = grid_plan(h, h->grid_stream ? h->grid_stream : h->stream);
  if (rc) return rc;
  *n_substeps = h->last_substeps;
  h->plan_open = true;
  h->plan_done = 0;
  return EPH_B200_OK;
}

int eph_b200_grid_substep(eph_b200_handle *h, int z_begin, int z_end) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->plan_open || h->plan_done >= h->last_substeps) return fail(h, EPH_B200_ERR_ARG, "grid_substep: no sub-step of a planned solve is due");
  if (z_begin < 0 || z_end > h->nz || z_begin >= z_end) return fail(h, EPH_B200_ERR_ARG, "grid_substep: bad plane range [%d, %d) of %d", z_begin, z_end, h->nz);
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = h->grid_stream ? h->grid_stream : h->stream;
  const bool last = h->plan_done + 1 == h->last_substeps;
  int rc = grid_substep(h, st, z_begin, z_end, last);
  if (rc) return rc;
  if (last) {   // the source term is consumed: the kernel cleared this rank's slab, the other planes are cleared here
    double *src = h->dT_e_ext ? h->dT_e_ext : h->dT_e.p;
    const size_t plane = (size_t)h->nx * h->ny;
    if (z_begin > 0) EPH_CUDA(h, cudaMemsetAsync(src, 0, (size_t)z_begin * plane * sizeof(double), st));
    if (z_end < h->nz) EPH_CUDA(h, cudaMemsetAsync(src + (size_t)z_end * plane, 0, (size_t)(h->nz - z_end) * plane * sizeof(double), st));
  }
  ++h->plan_done;
  return EPH_B200_OK;
}

int eph_b200_end_of_step_end_external(eph_b200_handle *h, double *E_local) {
  if (h && h->plan_open && h->plan_done != h->last_substeps)
    return fail(h, EPH_B200_ERR_ARG, "end_of_step_end_external: %d of %d planned sub-steps run", h->plan_done, h->last_substeps);
  return end_of_step_finish(h, E_local, false);
}

int eph_b200_grid_device_ptr(eph_b200_handle *h, int which, double **ptr) {
  if (!h || !ptr) return EPH_B200_ERR_ARG;
  if (!h->grid_set) return fail(h, EPH_B200_ERR_ARG, "grid_device_ptr: no grid");
  *ptr = grid_field(h, which);
  if (!*ptr) return fail(h, EPH_B200_ERR_ARG, "grid_device_ptr: bad field id %d", which);
  return EPH_B200_OK;
}

int eph_b200_end_of_step(eph_b200_handle *h, const double *x, const double *v, double *E_local, int memspace) {
  int rc = eph_b200_end_of_step_begin(h, x, v, memspace);
  if (rc) return rc;
  // several ranks with the engine's own transport: the source term is summed over ranks (eph_fdm.h:481) and every rank
  // solves the whole grid, or its slab of it (eph_b200_set_grid_sharding)
  if (h->comm && h->comm_size > 1) return eph_b200_reduce_and_solve(h, E_local);
  return eph_b200_end_of_step_end(h, E_local);
}

int eph_b200_set_grid_stream(eph_b200_handle *h, void *stream) {
  if (!h) return EPH_B200_ERR_ARG;
  cudaSetDevice(h->cfg.device);
  join_grid_stream(h);
  if (h->grid_stream) cudaStreamSynchronize(h->grid_stream);
  h->grid_stream = static_cast<cudaStream_t>(stream);
  return EPH_B200_OK;
}

int eph_b200_set_comm_stream(eph_b200_handle *h, void *stream) {
  if (!h) return EPH_B200_ERR_ARG;
  cudaSetDevice(h->cfg.device);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  cudaStreamSynchronize(h->stream);
  h->comm_stream = static_cast<cudaStream_t>(stream);
  h->unpack_pending = false;
  h->boundary_recorded = false;
  return EPH_B200_OK;
}

int eph_b200_set_boundary_atoms(eph_b200_handle *h, int n, const int *index, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "set_boundary_atoms: call set_atoms first");
  if (n < 0 || (n > 0 && !index)) return fail(h, EPH_B200_ERR_ARG, "set_boundary_atoms: bad arguments");
  cudaSetDevice(h->cfg.device);
  h->split_ready = false;
  h->n_first = h->n_rest = 0;
  const int nl = h->nlocal;
  if (nl == 0) return EPH_B200_OK;
  const int tile_atoms = 32 / h->lanes;
  const int ntiles = (nl + tile_atoms - 1) / tile_atoms;
  const int *didx = nullptr;
  int rc;
  if (n > 0 && (rc = stage_in(h, h->comm_idx, index, (size_t)n, memspace, &didx))) return rc;
  EPH_CUDA(h, h->tile_flag.reserve((size_t)ntiles + 1)); EPH_CUDA(h, h->tile_scan.reserve((size_t)ntiles + 1));
  EPH_CUDA(h, h->work_all.reserve(ntiles));
  if (!h->done_counter.p) {
    EPH_CUDA(h, h->done_counter.reserve(1));
    EPH_CUDA(h, cudaMemsetAsync(h->done_counter.p, 0, sizeof(unsigned), h->stream));
    h->boundary_target = 0;
  }
  EPH_CUDA(h, cudaMemsetAsync(h->tile_flag.p, 0, ((size_t)ntiles + 1) * sizeof(int), h->stream));
  if (n > 0) {
    tile_mark_kernel<<<blocks_for(n, 256), 256, 0, h->stream>>>(n, didx, tile_atoms, nl, h->tile_flag.p);
    EPH_LAUNCH_CHECK(h);
  }
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, h->tile_flag.p, h->tile_scan.p, ntiles + 1, h->stream);
  EPH_CUDA(h, h->nb_tmp.reserve(scan_bytes));
  EPH_CUDA(h, cub::DeviceScan::ExclusiveSum(h->nb_tmp.p, scan_bytes, h->tile_flag.p, h->tile_scan.p, ntiles + 1, h->stream));
  ++h->launches;
  tile_split_kernel<<<blocks_for(ntiles, 256), 256, 0, h->stream>>>(ntiles, h->tile_flag.p, h->tile_scan.p, h->work_all.p);
  EPH_LAUNCH_CHECK(h);
  int first = 0;
  EPH_CUDA(h, cudaMemcpyAsync(&first, h->tile_scan.p + ntiles, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_first = first;
  h->n_rest = ntiles - first;
  h->split_ready = true;
  return EPH_B200_OK;
}

int eph_b200_bind_grid_source(eph_b200_handle *h, double *dT_e_dev) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->grid_set) return fail(h, EPH_B200_ERR_ARG, "bind_grid_source: set_grid not called");
  h->dT_e_ext = dT_e_dev;   // nullptr: back to the library-owned array
  return EPH_B200_OK;
}

int eph_b200_initial_integrate(eph_b200_handle *h, double *x, double *v, const double *f, const double *mass_by_type,
                               double dtv, double dtf, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (h->cfg.flags & EPH_B200_NOINT) return EPH_B200_OK;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "initial_integrate: set_atoms not called");
  if (!x || !v || !f || !mass_by_type) return fail(h, EPH_B200_ERR_ARG, "initial_integrate: null argument");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  if (nl == 0) return EPH_B200_OK;
  { const int rc_m = upload_masses(h, mass_by_type); if (rc_m) return rc_m; }
  double *dx = x, *dv = v;
  const double *df = f;
  if (memspace != EPH_B200_DEVICE) {
    EPH_CUDA(h, h->x.reserve(3 * (size_t)nl)); EPH_CUDA(h, h->v.reserve(3 * (size_t)nl)); EPH_CUDA(h, h->f.reserve(3 * (size_t)nl));
    EPH_CUDA(h, cudaMemcpyAsync(h->x.p, x, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(h->v.p, v, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(h->f.p, f, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    dx = h->x.p; dv = h->v.p; df = h->f.p;
  }
  {
    KernelTimer kt(h, "initial_integrate");
    integrate_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(nl, dx, dv, df, h->type.p, h->mask.p, h->mass.p, h->cfg.groupbit, dtv, dtf, 1);
  }
  EPH_LAUNCH_CHECK(h);
  if (memspace != EPH_B200_DEVICE) {
    EPH_CUDA(h, cudaMemcpyAsync(x, h->x.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(v, h->v.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

int eph_b200_final_integrate(eph_b200_handle *h, double *v, const double *f, const double *mass_by_type, double dtf,
                             int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (h->cfg.flags & EPH_B200_NOINT) return EPH_B200_OK;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "final_integrate: set_atoms not called");
  if (!v || !f || !mass_by_type) return fail(h, EPH_B200_ERR_ARG, "final_integrate: null argument");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  if (nl == 0) return EPH_B200_OK;
  { const int rc_m = upload_masses(h, mass_by_type); if (rc_m) return rc_m; }
  double *dv = v;
  const double *df = f;
  if (memspace != EPH_B200_DEVICE) {
    EPH_CUDA(h, h->v.reserve(3 * (size_t)nl)); EPH_CUDA(h, h->f.reserve(3 * (size_t)nl));
    EPH_CUDA(h, cudaMemcpyAsync(h->v.p, v, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(h->f.p, f, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    dv = h->v.p; df = h->f.p;
  }
  {
    KernelTimer kt(h, "final_integrate");
    integrate_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(nl, nullptr, dv, df, h->type.p, h->mask.p, h->mass.p, h->cfg.groupbit, 0.0, dtf, 0);
  }
  EPH_LAUNCH_CHECK(h);
  if (memspace != EPH_B200_DEVICE) {
    EPH_CUDA(h, cudaMemcpyAsync(v, h->v.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

int eph_b200_get_peratom(eph_b200_handle *h, double *array8, int memspace) {
  if (!h || !array8) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "get_peratom: set_atoms not called");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  if (nl == 0) return EPH_B200_OK;
  // the 8 columns are not written every step (64 bytes per atom nobody reads on the device): they are formed here
  // from rho_i, f_EPH and f_RNG of the last step; before the first end_of_step they are zero like the reference's
  double *dst = memspace == EPH_B200_DEVICE ? array8 : h->array8.p;
  if (!h->peratom_valid) {
    EPH_CUDA(h, cudaMemsetAsync(dst, 0, 8 * (size_t)nl * sizeof(double), h->stream));
  } else {
    KernelTimer kt(h, "peratom");
    peratom_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(nl, h->pos4.p, h->rho.p, h->f_eph.p, h->f_rng.p, h->beta_tab.p,
                                                              h->n_beta, h->inv_drho, h->rho_cut, dst);
    EPH_LAUNCH_CHECK(h);
  }
  if (memspace != EPH_B200_DEVICE)
    EPH_CUDA(h, cudaMemcpyAsync(array8, h->array8.p, 8 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPH_B200_OK;
}

int eph_b200_get_probe(eph_b200_handle *h, int which, double *out) {
  if (!h || !out) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "get_probe: set_atoms not called");
  cudaSetDevice(h->cfg.device);
  const size_t nl = h->nlocal, nt = nl + h->nghost;
  const double *src = nullptr;
  size_t n = 0;
  switch (which) {
    case 0: src = h->rho.p; n = nt; break;
    case 1: src = h->w.p; n = 3 * nl; break;
    case 2: src = h->xi.p; n = 3 * nl; break;
    case 3: src = h->f_eph.p; n = 3 * nl; break;
    case 4: src = h->f_rng.p; n = 3 * nl; break;
    default: return fail(h, EPH_B200_ERR_ARG, "get_probe: bad id %d", which);
  }
  EPH_CUDA(h, cudaMemcpyAsync(out, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPH_B200_OK;
}

// Host transport (LAMMPS' own MPI comm): payload of `state` between device-resident arrays and a host buffer.
// RHO: rho (1 double); WI: the pair sums W of the density pass (3 doubles; the receiver forms w = s W itself);
// XI: the injected Gaussians (3 doubles).  Valid between post_force_begin and post_force_end.
static int forward_source(eph_b200_handle *h, int state, double **base, int *width, int *stride) {
  switch (state) {
    case EPH_B200_STATE_RHO: *base = h->rho.p; *width = 1; *stride = 1; return 0;
    case EPH_B200_STATE_WI: *base = reinterpret_cast<double *>(h->W4.p); *width = 3; *stride = 4; return 0;
    case EPH_B200_STATE_XI: *base = h->xi.p; *width = 3; *stride = 3; return 0;
    default: return -1;
  }
}

int eph_b200_pack_forward(eph_b200_handle *h, int state, int n, const int *list, double *buf) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->pf_open) return fail(h, EPH_B200_ERR_ARG, "pack_forward: only valid between post_force_begin and post_force_end");
  double *base; int width, stride;
  if (forward_source(h, state, &base, &width, &stride)) return fail(h, EPH_B200_ERR_ARG, "pack_forward: bad state %d", state);
  if (n <= 0) return EPH_B200_OK;
  if (!list || !buf) return fail(h, EPH_B200_ERR_ARG, "pack_forward: null list or buffer");
  cudaSetDevice(h->cfg.device);
  EPH_CUDA(h, h->comm_idx.reserve(n)); EPH_CUDA(h, h->comm_buf.reserve((size_t)n * width));
  EPH_CUDA(h, cudaMemcpyAsync(h->comm_idx.p, list, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  pack_forward_kernel<<<blocks_for((long long)n * width, 256), 256, 0, h->stream>>>(n, h->comm_idx.p, base, width, stride, h->comm_buf.p);
  EPH_LAUNCH_CHECK(h);
  EPH_CUDA(h, cudaMemcpyAsync(buf, h->comm_buf.p, (size_t)n * width * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return n * width;
}

int eph_b200_unpack_forward(eph_b200_handle *h, int state, int n, int first, const double *buf) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->pf_open) return fail(h, EPH_B200_ERR_ARG, "unpack_forward: only valid between post_force_begin and post_force_end");
  double *base; int width, stride;
  if (forward_source(h, state, &base, &width, &stride)) return fail(h, EPH_B200_ERR_ARG, "unpack_forward: bad state %d", state);
  if (n <= 0) return EPH_B200_OK;
  if (!buf || first < 0 || first + n > h->nlocal + h->nghost) return fail(h, EPH_B200_ERR_ARG, "unpack_forward: bad range");
  cudaSetDevice(h->cfg.device);
  EPH_CUDA(h, h->comm_buf.reserve((size_t)n * width));
  EPH_CUDA(h, cudaMemcpyAsync(h->comm_buf.p, buf, (size_t)n * width * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  unpack_forward_kernel<<<blocks_for((long long)n * width, 256), 256, 0, h->stream>>>(n, first, base, width, stride, h->comm_buf.p);
  EPH_LAUNCH_CHECK(h);
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));   // the caller may reuse its buffer
  return EPH_B200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Multi-rank data plane inside the engine: NCCL over NVLink carries the ghost exchange of post_force
// (reference: the forward comms RHO, WI, XI, fix_eph.cpp:743-744, :863-871), the all-reduce of the grid source term
// (eph_fdm.h:481) and the halo planes of the sharded grid solve (which replaces rank-0 solve + MPI_Bcast, eph_fdm.h:271, :490).
// ---------------------------------------------------------------------------
namespace {

#define EPH_NCCL(h, call)                                                                                         \
  do {                                                                                                            \
    const int r_ = (call);                                                                                        \
    if (r_ != kNcclSuccess)                                                                                       \
      return fail(h, EPH_B200_ERR_COMM, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

// rows of `width` doubles from a receive buffer into the listed slots of a [n][stride] array
__global__ void scatter_rows_kernel(int n, const int *__restrict__ slot, double *__restrict__ dst, int width, int stride,
                                    const double *__restrict__ buf, int nrows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const int k = t / width, c = t - k * width;
  const int a = slot[k];
  if (a >= 0 && a < nrows) dst[(size_t)a * stride + c] = buf[t];
}

// Peer-memory exchange (eph_p2p.cuh): allocate and export this rank's window, learn everybody's handle through the
// communicator itself, map them, and agree -- all ranks or none -- that the exchanges go that way.  Failing to map a
// peer (another node, no peer access) is not an error: the NCCL send / receive path stays.
int p2p_setup(eph_b200_handle *h) {
  p2p_teardown(h);
  NcclApi &api = nccl_api();
  struct Record { cudaIpcMemHandle_t handle; unsigned long long bytes; int want; int pad; };
  const char *mode = std::getenv("EPH_B200_EXCHANGE");
  const char *mb = std::getenv("EPH_B200_P2P_WINDOW_MB");
  Record mine{};
  mine.bytes = (unsigned long long)(std::max(0.0625, mb ? std::atof(mb) : 256.0) * 1048576.0) / 4096 * 4096;
  mine.want = !(mode && std::strcmp(mode, "nccl") == 0) && h->comm_size > 1 && h->comm_size <= kP2PMaxRanks;
  if (mine.want) {
    void *w = nullptr;
#ifdef EPHA_HOST_EMULATION
    if (emul_ipc_malloc(&w, mine.bytes) != cudaSuccess ||   // tests/emul: a shared-memory object the other ranks can map
#else
    if (cudaMalloc(&w, mine.bytes) != cudaSuccess ||
#endif cudaMemset(w, 0, kP2PHeaderBytes) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.handle, w) != cudaSuccess) {
      cudaGetLastError();
      if (w) cudaFree(w);
      mine.want = 0;
    } else {
      h->p2p_window = static_cast<char *>(w);
      h->p2p_window_bytes = mine.bytes;
    }
  }
  const int n = h->comm_size;
  DevBuf<unsigned char> rec;
  EPH_CUDA(h, rec.reserve(sizeof(Record) * (size_t)n));
  EPH_CUDA(h, cudaMemcpy(rec.p + sizeof(Record) * (size_t)h->comm_rank, &mine, sizeof(Record), cudaMemcpyHostToDevice));
  EPH_CUDA(h, cudaDeviceSynchronize());   // the window's cleared header and the record precede the collective on h->stream
  EPH_NCCL(h, api.AllGather(rec.p + sizeof(Record) * (size_t)h->comm_rank, rec.p, sizeof(Record), kNcclInt8, h->comm, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  std::vector<Record> all((size_t)n);
  EPH_CUDA(h, cudaMemcpy(all.data(), rec.p, sizeof(Record) * (size_t)n, cudaMemcpyDeviceToHost));
  rec.release();
  int ok = mine.want;
  for (int r = 0; r < n; ++r)
    if (!all[r].want || all[r].bytes != mine.bytes) ok = 0;
  h->p2p_peer_window.assign((size_t)n, nullptr);
  if (ok) {
    for (int r = 0; r < n && ok; ++r) {
      if (r == h->comm_rank) { h->p2p_peer_window[r] = h->p2p_window; continue; }
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
      else h->p2p_peer_window[r] = static_cast<char *>(ptr);
    }
  }
  EPH_CUDA(h, h->p2p_scratch.reserve(2));
  EPH_CUDA(h, h->p2p_done.reserve(2));
  EPH_CUDA(h, cudaMemset(h->p2p_done.p, 0, 2 * sizeof(unsigned)));
  const double okd = ok;
  double sum = 0.0;
  EPH_CUDA(h, cudaMemcpy(h->p2p_scratch.p, &okd, sizeof(double), cudaMemcpyHostToDevice));
  EPH_CUDA(h, cudaDeviceSynchronize());
  EPH_NCCL(h, api.AllReduce(h->p2p_scratch.p, h->p2p_scratch.p, 1, kNcclFloat64, kNcclSum, h->comm, h->stream));
  EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  EPH_CUDA(h, cudaMemcpy(&sum, h->p2p_scratch.p, sizeof(double), cudaMemcpyDeviceToHost));
  if (sum == (double)n) {
    h->p2p_ok = true;
    h->p2p_region_bytes = (h->p2p_window_bytes - kP2PHeaderBytes) / (size_t)n / 512 * 512;
    h->p2p_epoch[0] = h->p2p_epoch[1] = 0;
  } else {
    p2p_teardown(h);
  }
  return EPH_B200_OK;
}

// z-planes [z0, z1) this rank advances in a sharded solve; false if the grid does not divide
bool grid_slab(const eph_b200_handle *h, int *z0, int *z1) {
  if (h->comm_size < 1 || h->nz % h->comm_size) return false;
  const int per = h->nz / h->comm_size;
  *z0 = h->comm_rank * per;
  *z1 = *z0 + per;
  return true;
}

}  // namespace

extern "C" {

int eph_b200_comm_get_id(void *id128) {
  if (!id128) return EPH_B200_ERR_ARG;
  NcclApi &api = nccl_api();
  if (!api.ok) { g_create_error = "eph_b200_comm_get_id: " + api.error; return EPH_B200_ERR_COMM; }
  NcclUniqueId id;
  const int r = api.GetUniqueId(&id);
  if (r != kNcclSuccess) { g_create_error = std::string("ncclGetUniqueId failed: ") + api.GetErrorString(r); return EPH_B200_ERR_COMM; }
  std::memcpy(id128, &id, sizeof id);
  return EPH_B200_OK;
}

int eph_b200_comm_init(eph_b200_handle *h, const void *id128, int rank, int nranks) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, EPH_B200_ERR_ARG, "comm_init: bad rank %d of %d or null id", rank, nranks);
  NcclApi &api = nccl_api();
  if (!api.ok) return fail(h, EPH_B200_ERR_COMM, "comm_init: %s (the multi-rank data plane needs NCCL; there is no host fall-back)", api.error.c_str());
  cudaSetDevice(h->cfg.device);
  if (h->comm) { api.CommDestroy(h->comm); h->comm = nullptr; }
  NcclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  EPH_NCCL(h, api.CommInitRank(&h->comm, nranks, id, rank));
  h->comm_rank = rank;
  h->comm_size = nranks;
  h->ghost_map_set = false;
  return p2p_setup(h);
}

int eph_b200_comm_transport(const eph_b200_handle *h) {
  if (!h || !h->comm || h->comm_size < 2) return 0;
  return h->p2p_ok ? 2 : 1;
}

int eph_b200_set_ghost_map(eph_b200_handle *h, int npeers, const int *peer_rank, const int *send_count, const int *send_index,
                           const int *recv_count, const int *recv_slot) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: call set_atoms first");
  if (npeers < 0 || (npeers > 0 && (!peer_rank || !send_count || !recv_count))) return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: bad arguments");
  cudaSetDevice(h->cfg.device);
  long long ns = 0, nr = 0;
  for (int p = 0; p < npeers; ++p) {
    if (peer_rank[p] < 0 || peer_rank[p] >= h->comm_size || peer_rank[p] == h->comm_rank || send_count[p] < 0 || recv_count[p] < 0)
      return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: peer %d: rank %d, %d to send, %d to receive", p, peer_rank[p], send_count[p], recv_count[p]);
    ns += send_count[p]; nr += recv_count[p];
  }
  if ((ns > 0 && !send_index) || (nr > 0 && !recv_slot)) return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: null index lists");
  for (long long k = 0; k < ns; ++k)
    if (send_index[k] < 0 || send_index[k] >= h->nlocal) return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: send index %d is not an owned atom", send_index[k]);
  for (long long k = 0; k < nr; ++k)
    if (recv_slot[k] < h->nlocal || recv_slot[k] >= h->nlocal + h->nghost) return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: receive slot %d is not a ghost", recv_slot[k]);
  h->gm_peer.assign(peer_rank, peer_rank + npeers);
  h->gm_send_count.assign(send_count, send_count + npeers);
  h->gm_recv_count.assign(recv_count, recv_count + npeers);
  h->gm_nsend = (int)ns; h->gm_nrecv = (int)nr;
  EPH_CUDA(h, h->gm_send_idx.reserve(std::max<size_t>(ns, 1))); EPH_CUDA(h, h->gm_recv_slot.reserve(std::max<size_t>(nr, 1)));
  EPH_CUDA(h, h->gm_send_buf.reserve(4 * std::max<size_t>(ns, 1))); EPH_CUDA(h, h->gm_recv_buf.reserve(4 * std::max<size_t>(nr, 1)));
  EPH_CUDA(h, h->gm_send_xi.reserve(3 * std::max<size_t>(ns, 1))); EPH_CUDA(h, h->gm_recv_xi.reserve(3 * std::max<size_t>(nr, 1)));
  if (ns) EPH_CUDA(h, cudaMemcpyAsync(h->gm_send_idx.p, send_index, ns * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (nr) EPH_CUDA(h, cudaMemcpyAsync(h->gm_recv_slot.p, recv_slot, nr * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (!h->p2p_ok) EPH_CUDA(h, cudaStreamSynchronize(h->stream));   // (with peer memory the synchronisation below covers the copies)
  if (h->p2p_ok) {
    P2PMap &m = h->p2p_map;
    m = P2PMap{};
    const bool too_many = npeers > kP2PMaxPeers;
    m.n = too_many ? 0 : npeers; m.my_rank = h->comm_rank; m.local = h->p2p_window; m.region_bytes = h->p2p_region_bytes;
    const char *tmo = std::getenv("EPH_B200_P2P_TIMEOUT_MS");
    m.timeout_ns = tmo && std::atof(tmo) > 0.0 ? (unsigned long long)(std::atof(tmo) * 1e6) : kP2PTimeoutNs;
    size_t worst = 0;
    int worst_rows = 0, worst_rank = -1;
    for (int p = 0; p < m.n; ++p) {
      m.rank[p] = peer_rank[p];
      m.send_off[p + 1] = m.send_off[p] + send_count[p];
      m.recv_off[p + 1] = m.recv_off[p] + recv_count[p];
      m.remote[p] = h->p2p_peer_window[peer_rank[p]];
      const size_t need = p2p_half_bytes_needed(std::max(send_count[p], recv_count[p]));
      if (need > worst) { worst = need; worst_rows = std::max(send_count[p], recv_count[p]); worst_rank = peer_rank[p]; }
    }
    // Re-registration is collective: nobody writes rows of the new map into a window whose owner may still be reading
    // rows of the old one (with an unchanged set of peers the exchanges themselves guarantee that).  The same
    // all-reduce carries "my rows do not fit", so that every rank fails together instead of one leaving the others
    // waiting in the next collective.
    const double mine = (too_many || worst > m.region_bytes / 2) ? 1.0 : 0.0;
    double failed = 0.0;
    EPH_CUDA(h, cudaMemcpyAsync(h->p2p_scratch.p + 1, &mine, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EPH_NCCL(h, nccl_api().AllReduce(h->p2p_scratch.p + 1, h->p2p_scratch.p + 1, 1, kNcclFloat64, kNcclSum, h->comm, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(&failed, h->p2p_scratch.p + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
    if (too_many)
      return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: %d peers, the peer-memory exchange takes %d (EPH_B200_EXCHANGE=nccl selects send/receive)", npeers, kP2PMaxPeers);
    if (mine > 0.0)
      return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: %d ghost rows exchanged with rank %d need %zu bytes of a peer-memory window, a rank's share "
                  "is %zu (raise EPH_B200_P2P_WINDOW_MB on all ranks, or EPH_B200_EXCHANGE=nccl)", worst_rows, worst_rank, worst, m.region_bytes / 2);
    if (failed > 0.0)
      return fail(h, EPH_B200_ERR_ARG, "set_ghost_map: the ghost rows of %d other rank(s) do not fit their peer-memory windows "
                  "(raise EPH_B200_P2P_WINDOW_MB on all ranks, or EPH_B200_EXCHANGE=nccl)", (int)failed);
  }
  h->ghost_map_set = true;
  // with a communication stream registered the density pass sweeps the tiles these atoms live in first
  if (h->comm_stream && ns > 0) return eph_b200_set_boundary_atoms(h, (int)ns, h->gm_send_idx.p, EPH_B200_DEVICE);
  return EPH_B200_OK;
}

// The one ghost exchange of a step: {rho, W} of the owned atoms other ranks hold as ghosts (and the injected xi, when the
// caller supplies the Gaussians instead of the tag-keyed stream), one grouped ncclSend / ncclRecv per peer.
int eph_b200_exchange_ghosts(eph_b200_handle *h) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->pf_open) return fail(h, EPH_B200_ERR_ARG, "exchange_ghosts: only valid between post_force_begin and post_force_end");
  if (!h->comm) return fail(h, EPH_B200_ERR_ARG, "exchange_ghosts: eph_b200_comm_init has not been called");
  if (!h->ghost_map_set) return fail(h, EPH_B200_ERR_ARG, "exchange_ghosts: eph_b200_set_ghost_map must follow every set_atoms");
  cudaSetDevice(h->cfg.device);
  NcclApi &api = nccl_api();
  const bool with_xi = h->pf_xi != nullptr && (h->cfg.flags & EPH_B200_RANDOM);
  cudaStream_t st = h->comm_stream ? h->comm_stream : h->stream;
  int rc;
  if (h->p2p_ok) {
    // peer memory: the sending kernel stores the rows into the receivers' windows and raises its flag there
    if ((rc = exchange_stream(h, &st))) return rc;
    const P2PMap &m = h->p2p_map;
    if (m.n > 0) {
      const unsigned long long epoch = ++h->p2p_epoch[0];
      const int cap = 2 * h->sm_count;   // all blocks resident: the receiving blocks poll
      {
        KernelTimer kt(h, "pack_payload", st);
        p2p_send_payload_kernel<<<std::min(cap, std::max(1, blocks_for(h->gm_nsend, 256))), 256, 0, st>>>(
            m, h->gm_nsend, h->gm_send_idx.p, h->rho.p, h->W4.p, with_xi ? h->xi.p : nullptr, epoch, h->p2p_done.p);
      }
      EPH_LAUNCH_CHECK(h);
      {
        KernelTimer kt(h, "unpack_payload", st);
        p2p_recv_payload_kernel<<<std::min(cap, std::max(1, blocks_for(h->gm_nrecv, 256))), 256, 0, st>>>(
            m, h->gm_nrecv, h->gm_recv_slot.p, h->rho.p, h->W4.p, with_xi ? h->xi.p : nullptr, h->nlocal + h->nghost, epoch, h->d_status.p);
      }
      EPH_LAUNCH_CHECK(h);
    }
    if (h->comm_stream) {   // post_force_end waits for the ghosts' {rho, W}
      EPH_CUDA(h, cudaEventRecord(h->ev_unpacked, st));
      h->unpack_pending = true;
    }
    return EPH_B200_OK;
  }
  if (h->gm_nsend) {
    if ((rc = eph_b200_pack_ghost_payload(h, h->gm_nsend, h->gm_send_idx.p, h->gm_send_buf.p))) return rc;
    if (with_xi) {
      pack_forward_kernel<<<blocks_for(3LL * h->gm_nsend, 256), 256, 0, st>>>(h->gm_nsend, h->gm_send_idx.p, h->xi.p, 3, 3, h->gm_send_xi.p);
      EPH_LAUNCH_CHECK(h);
    }
  } else if (h->comm_stream && h->boundary_recorded) {
    EPH_CUDA(h, cudaStreamWaitEvent(st, h->ev_boundary, 0));
  }
  {
    KernelTimer kt(h, "ghost_exchange", st);
    EPH_NCCL(h, api.GroupStart());
    size_t so = 0, ro = 0;
    for (size_t p = 0; p < h->gm_peer.size(); ++p) {
      const size_t sc = h->gm_send_count[p], rcn = h->gm_recv_count[p];
      if (sc) {
        EPH_NCCL(h, api.Send(h->gm_send_buf.p + 4 * so, 4 * sc, kNcclFloat64, h->gm_peer[p], h->comm, st));
        if (with_xi) EPH_NCCL(h, api.Send(h->gm_send_xi.p + 3 * so, 3 * sc, kNcclFloat64, h->gm_peer[p], h->comm, st));
      }
      if (rcn) {
        EPH_NCCL(h, api.Recv(h->gm_recv_buf.p + 4 * ro, 4 * rcn, kNcclFloat64, h->gm_peer[p], h->comm, st));
        if (with_xi) EPH_NCCL(h, api.Recv(h->gm_recv_xi.p + 3 * ro, 3 * rcn, kNcclFloat64, h->gm_peer[p], h->comm, st));
      }
      so += sc; ro += rcn;
    }
    EPH_NCCL(h, api.GroupEnd());
  }
  if (h->gm_nrecv) {
    if (with_xi) {   // the XI forward comm of the reference (fix_eph.cpp:863-864): the ghost's slot of xi
      scatter_rows_kernel<<<blocks_for(3LL * h->gm_nrecv, 256), 256, 0, st>>>(h->gm_nrecv, h->gm_recv_slot.p, h->xi.p, 3, 3, h->gm_recv_xi.p,
                                                                              h->nlocal + h->nghost);
      EPH_LAUNCH_CHECK(h);
    }
    if ((rc = eph_b200_unpack_ghost_payload(h, h->gm_nrecv, h->gm_recv_slot.p, h->gm_recv_buf.p))) return rc;
  } else if (h->comm_stream) {
    EPH_CUDA(h, cudaEventRecord(h->ev_unpacked, st));
    h->unpack_pending = true;
  }
  return EPH_B200_OK;
}

int eph_b200_set_grid_sharding(eph_b200_handle *h, int on) {
  if (!h) return EPH_B200_ERR_ARG;
  h->grid_sharded = on != 0;
  return EPH_B200_OK;
}

// Second half of end_of_step on several ranks: ncclAllReduce of the source term (the reference's MPI_Allreduce,
// eph_fdm.h:481), then the solve -- the whole grid on every rank, or with sharding this rank's z-slab with one plane pair
// exchanged between sub-steps (periodic in z, written in place at their global position) and one all-gather of the
// slabs at the end, which replaces the reference's MPI_Bcast (eph_fdm.h:490).
int eph_b200_reduce_and_solve(eph_b200_handle *h, double *E_local) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->eos_open) return fail(h, EPH_B200_ERR_ARG, "reduce_and_solve without end_of_step_begin");
  if (!h->comm) return fail(h, EPH_B200_ERR_ARG, "reduce_and_solve: eph_b200_comm_init has not been called");
  cudaSetDevice(h->cfg.device);
  NcclApi &api = nccl_api();
  cudaStream_t st = h->grid_stream ? h->grid_stream : h->stream;
  double *src = h->dT_e_ext ? h->dT_e_ext : h->dT_e.p;
  int z0 = 0, z1 = 0;
  const bool slabs = h->grid_sharded && (h->cfg.flags & EPH_B200_FDM) && grid_slab(h, &z0, &z1);
  if (!slabs) {
    {
      KernelTimer kt(h, "source_allreduce", st);
      EPH_NCCL(h, api.AllReduce(src, src, (size_t)h->ncell, kNcclFloat64, kNcclSum, h->comm, st));
    }
    return eph_b200_end_of_step_end(h, E_local);
  }
  {
    // a slab solve needs the summed source term of its own planes only: reduce-scatter (in place: the slab sits where it
    // belongs), half the traffic of the all-reduce whose other half the all-gather of T_e below makes up for
    KernelTimer kt(h, "source_reduce_scatter", st);
    const size_t slab = (size_t)(z1 - z0) * h->nx * h->ny;
    EPH_NCCL(h, api.ReduceScatter(src, src + (size_t)z0 * h->nx * h->ny, slab, kNcclFloat64, kNcclSum, h->comm, st));
  }
  int rc = grid_plan(h, st);
  if (rc) return rc;
  const int n = h->last_substeps;
  const size_t plane = (size_t)h->nx * h->ny;
  const int prev = (h->comm_rank + h->comm_size - 1) % h->comm_size, next = (h->comm_rank + 1) % h->comm_size;
  for (int sstep = 0; sstep < n; ++sstep) {
    if ((rc = grid_substep(h, st, z0, z1, sstep + 1 == n))) return rc;
    if (sstep + 1 == n) break;
    // halo planes for the next sub-step.  Posting order matters when prev == next (two ranks): sends go "bottom plane to
    // prev, top plane to next", receives "from next into the plane above the slab, from prev into the plane below"
    double *T = h->T[h->cur].p;
    const int zlo = (z0 + h->nz - 1) % h->nz, zhi = z1 % h->nz;
    KernelTimer kt(h, "grid_halo", st);
    EPH_NCCL(h, api.GroupStart());
    EPH_NCCL(h, api.Send(T + (size_t)z0 * plane, plane, kNcclFloat64, prev, h->comm, st));
    EPH_NCCL(h, api.Send(T + (size_t)(z1 - 1) * plane, plane, kNcclFloat64, next, h->comm, st));
    EPH_NCCL(h, api.Recv(T + (size_t)zhi * plane, plane, kNcclFloat64, next, h->comm, st));
    EPH_NCCL(h, api.Recv(T + (size_t)zlo * plane, plane, kNcclFloat64, prev, h->comm, st));
    EPH_NCCL(h, api.GroupEnd());
  }
  if (n > 0) {
    // the kernel cleared the source term on this rank's slab only
    if (z0 > 0) EPH_CUDA(h, cudaMemsetAsync(src, 0, (size_t)z0 * plane * sizeof(double), st));
    if (z1 < h->nz) EPH_CUDA(h, cudaMemsetAsync(src + (size_t)z1 * plane, 0, (size_t)(h->nz - z1) * plane * sizeof(double), st));
    double *T = h->T[h->cur].p;
    KernelTimer kt(h, "grid_allgather", st);
    EPH_NCCL(h, api.AllGather(T + (size_t)z0 * plane, T, (size_t)(z1 - z0) * plane, kNcclFloat64, h->comm, st));   // in place
  }
  return end_of_step_finish(h, E_local, false);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Ghost atoms following their owners on the device, and device-resident integration
// ---------------------------------------------------------------------------
namespace {

__global__ void add_rows_kernel(long long n, double *__restrict__ dst, const double *__restrict__ src) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) dst[t] += src[t];
}

// own periodic images: x_ghost = x_owner + shift, v_ghost = v_owner; record = 1 stores the shifts instead
__global__ void ghost_images_kernel(int nlocal, int nghost, const int *__restrict__ owner, double *__restrict__ x, double *__restrict__ v,
                                    double *__restrict__ shift, int record) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int o = owner[g];
  if (o < 0) return;   // owned by another rank: filled from the exchange
  const size_t a = 3 * (size_t)(nlocal + g), b = 3 * (size_t)o, c = 3 * (size_t)g;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (record) shift[c + d] = x[a + d] - x[b + d];
    else { x[a + d] = x[b + d] + shift[c + d]; v[a + d] = v[b + d]; }
  }
}
__global__ void pack_xv_kernel(int n, const int *__restrict__ index, const double *__restrict__ x, const double *__restrict__ v, double *__restrict__ buf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const size_t a = 3 * (size_t)index[t];
#pragma unroll
  for (int d = 0; d < 3; ++d) { buf[6 * (size_t)t + d] = x[a + d]; buf[6 * (size_t)t + 3 + d] = v[a + d]; }
}
__global__ void unpack_xv_kernel(int n, const int *__restrict__ slot, int nlocal, double *__restrict__ x, double *__restrict__ v, const double *__restrict__ buf,
                                 double *__restrict__ shift, int record) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const size_t a = 3 * (size_t)slot[t], c = 3 * (size_t)(slot[t] - nlocal);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (record) shift[c + d] = x[a + d] - buf[6 * (size_t)t + d];
    else { x[a + d] = buf[6 * (size_t)t + d] + shift[c + d]; v[a + d] = buf[6 * (size_t)t + 3 + d]; }
  }
}

}  // namespace

extern "C" {

// What LAMMPS' Comm::forward_comm() does for positions and velocities (comm->ghost_velocity, fix_eph.cpp:82), for callers
// whose x and v live on the device: images of this rank's own atoms are shifted copies, ghosts owned by other ranks
// arrive over the engine's NCCL ghost map.  The first call after set_atoms (when the ghost coordinates are still the
// ones LAMMPS produced) records the image shifts, later calls apply them.  x, v: DEVICE arrays [nlocal + nghost][3].
int eph_b200_refresh_ghosts(eph_b200_handle *h, double *x, double *v) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "refresh_ghosts: set_atoms not called");
  if (!x || !v) return fail(h, EPH_B200_ERR_ARG, "refresh_ghosts: null x or v");
  const int nl = h->nlocal, ng = h->nghost;
  const bool remote = h->comm && h->comm_size > 1;
  if (ng == 0 && !(remote && h->ghost_map_set && (h->gm_nsend || h->gm_nrecv || h->p2p_map.n > 0))) return EPH_B200_OK;
  if (ng > 0 && !h->has_owner) return fail(h, EPH_B200_ERR_ARG, "refresh_ghosts: set_atoms was given no ghost owners");
  if (remote && !h->ghost_map_set) return fail(h, EPH_B200_ERR_ARG, "refresh_ghosts: eph_b200_set_ghost_map must follow every set_atoms");
  cudaSetDevice(h->cfg.device);
  const int record = h->gshift_valid ? 0 : 1;
  EPH_CUDA(h, h->gshift.reserve(3 * (size_t)std::max(ng, 1)));
  if (remote && h->p2p_ok) {
    const P2PMap &m = h->p2p_map;
    if (m.n > 0) {
      const unsigned long long epoch = ++h->p2p_epoch[1];
      const int cap = 2 * h->sm_count;
      {
        KernelTimer kt(h, "refresh_send");
        p2p_send_xv_kernel<<<std::min(cap, std::max(1, blocks_for(3LL * h->gm_nsend, 256))), 256, 0, h->stream>>>(m, h->gm_nsend, h->gm_send_idx.p, x, v, epoch,
                                                                                                          h->p2p_done.p + 1);
      }
      EPH_LAUNCH_CHECK(h);
      {
        KernelTimer kt(h, "refresh_recv");
        p2p_recv_xv_kernel<<<std::min(cap, std::max(1, blocks_for(h->gm_nrecv, 256))), 256, 0, h->stream>>>(m, h->gm_nrecv, h->gm_recv_slot.p, nl, x, v,
                                                                                                          h->gshift.p, record, epoch, h->d_status.p);
      }
      EPH_LAUNCH_CHECK(h);
    }
  } else if (remote && (h->gm_nsend || h->gm_nrecv)) {
    NcclApi &api = nccl_api();
    EPH_CUDA(h, h->gm_send_xv.reserve(6 * std::max<size_t>(h->gm_nsend, 1))); EPH_CUDA(h, h->gm_recv_xv.reserve(6 * std::max<size_t>(h->gm_nrecv, 1)));
    if (h->gm_nsend) {
      pack_xv_kernel<<<blocks_for(h->gm_nsend, 256), 256, 0, h->stream>>>(h->gm_nsend, h->gm_send_idx.p, x, v, h->gm_send_xv.p);
      EPH_LAUNCH_CHECK(h);
    }
    EPH_NCCL(h, api.GroupStart());
    size_t so = 0, ro = 0;
    for (size_t p = 0; p < h->gm_peer.size(); ++p) {
      const size_t sc = h->gm_send_count[p], rcn = h->gm_recv_count[p];
      if (sc) EPH_NCCL(h, api.Send(h->gm_send_xv.p + 6 * so, 6 * sc, kNcclFloat64, h->gm_peer[p], h->comm, h->stream));
      if (rcn) EPH_NCCL(h, api.Recv(h->gm_recv_xv.p + 6 * ro, 6 * rcn, kNcclFloat64, h->gm_peer[p], h->comm, h->stream));
      so += sc; ro += rcn;
    }
    EPH_NCCL(h, api.GroupEnd());
    if (h->gm_nrecv) {
      unpack_xv_kernel<<<blocks_for(h->gm_nrecv, 256), 256, 0, h->stream>>>(h->gm_nrecv, h->gm_recv_slot.p, nl, x, v, h->gm_recv_xv.p, h->gshift.p, record);
      EPH_LAUNCH_CHECK(h);
    }
  }
  if (ng > 0) {
    KernelTimer kt(h, "ghost_images");
    ghost_images_kernel<<<blocks_for(ng, 256), 256, 0, h->stream>>>(nl, ng, h->owner.p, x, v, h->gshift.p, record);
  }
  EPH_LAUNCH_CHECK(h);
  h->gshift_valid = true;
  return EPH_B200_OK;
}

// ---- device-resident integration --------------------------------------------------------------------------------------
// x, v of all atoms (and f of the local ones) stay on the device between the fix hooks; per step only the pair forces go
// up and x (after the drift), f (after post_force) and v (after the second kick) come down for LAMMPS' own use.

int eph_b200_resident_upload(eph_b200_handle *h, const double *x, const double *v) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "resident_upload: set_atoms not called");
  if (h->cfg.flags & EPH_B200_NOINT) return fail(h, EPH_B200_ERR_ARG, "resident_upload: flag 8 (no integration) leaves the atoms to another integrator");
  if (!x || !v) return fail(h, EPH_B200_ERR_ARG, "resident_upload: null x or v");
  cudaSetDevice(h->cfg.device);
  const size_t nt = (size_t)h->nlocal + h->nghost, nl = h->nlocal;
  EPH_CUDA(h, h->res_x.reserve(3 * std::max<size_t>(nt, 1))); EPH_CUDA(h, h->res_v.reserve(3 * std::max<size_t>(nt, 1)));
  EPH_CUDA(h, h->res_f.reserve(3 * std::max<size_t>(nl, 1)));
  if (nt) {
    EPH_CUDA(h, cudaMemcpyAsync(h->res_x.p, x, 3 * nt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EPH_CUDA(h, cudaMemcpyAsync(h->res_v.p, v, 3 * nt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->resident = true;
  h->gshift_valid = false;
  if (h->nghost > 0) {   // the ghost coordinates just uploaded are LAMMPS' own: record the image shifts
    int rc = eph_b200_refresh_ghosts(h, h->res_x.p, h->res_v.p);
    if (rc) return rc;
  }
  return EPH_B200_OK;
}

// f: HOST total forces of the local atoms, needed only while the engine holds none itself (first step after an upload
// without a preceding resident_post_force; may be NULL otherwise).  x_out: HOST [nlocal][3], receives the new positions.
int eph_b200_resident_initial_integrate(eph_b200_handle *h, const double *f, const double *mass_by_type, double dtv, double dtf, double *x_out,
                                        long long start_post_force_step) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->resident) return fail(h, EPH_B200_ERR_ARG, "resident_initial_integrate: resident_upload not called since set_atoms");
  if (!mass_by_type) return fail(h, EPH_B200_ERR_ARG, "resident_initial_integrate: null masses");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  if (nl == 0) return EPH_B200_OK;
  if (!h->res_f_valid) {
    if (!f) return fail(h, EPH_B200_ERR_ARG, "resident_initial_integrate: the engine holds no forces yet and none were passed");
    EPH_CUDA(h, cudaMemcpyAsync(h->res_f.p, f, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->res_f_valid = true;
  }
  { const int rc_m = upload_masses(h, mass_by_type); if (rc_m) return rc_m; }
  integrate_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(nl, h->res_x.p, h->res_v.p, h->res_f.p, h->type.p, h->mask.p, h->mass.p, h->cfg.groupbit, dtv, dtf, 1);
  EPH_LAUNCH_CHECK(h);
  // x travels to the host on the copy stream while the main stream goes on: ghosts follow their owners and, if the caller
  // knows that LAMMPS will not re-neighbour in this step, the first half of post_force (records, density pass) starts
  // right away -- it needs x and v only, so it runs while the host computes its pair forces
  // whatever the copy stream does next (x down; this step's pair forces up, into the array the kick has just read) comes
  // after the kick
  EPH_CUDA(h, cudaEventRecord(h->f_event, h->stream));
  EPH_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->f_event, 0));
  if (x_out) EPH_CUDA(h, cudaMemcpyAsync(x_out, h->res_x.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
  int rc = eph_b200_refresh_ghosts(h, h->res_x.p, h->res_v.p);
  if (rc) return rc;
  if (start_post_force_step >= 0 && h->neigh_set) {
    if ((rc = eph_b200_post_force_begin(h, h->res_x.p, h->res_v.p, nullptr, start_post_force_step, EPH_B200_DEVICE))) return rc;
    h->res_pf_started = true;
  }
  if (x_out) EPH_CUDA(h, cudaStreamSynchronize(h->copy_stream));
  return EPH_B200_OK;
}

// f: HOST [nlocal][3] forces of the other force contributors (the pair style): go up behind the density pass; the engine
// keeps f + f_EPH (+ f_RNG) for the two kicks and copies it to f_out (HOST, may be f itself; NULL: not needed on the
// host this step).
int eph_b200_resident_post_force(eph_b200_handle *h, const double *f, double *f_out, const double *xi_inject, long long ntimestep) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->resident) return fail(h, EPH_B200_ERR_ARG, "resident_post_force: resident_upload not called since set_atoms");
  if (!f && h->nlocal > 0) return fail(h, EPH_B200_ERR_ARG, "resident_post_force: null f");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  const double *dxi = nullptr;
  int rc;
  if (xi_inject && nl > 0) {
    if ((rc = stage_in(h, h->xi_in, xi_inject, 3 * (size_t)nl, EPH_B200_HOST, &dxi))) return rc;
  }
  if (nl > 0) {   // the pair forces travel while the density pass runs
    EPH_CUDA(h, cudaMemcpyAsync(h->res_f.p, f, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    EPH_CUDA(h, cudaEventRecord(h->f_event, h->copy_stream));
  }
  if (h->res_pf_started && h->pf_open && !dxi && h->pf_step == ntimestep) {
    // resident_initial_integrate has started this step's density pass already
  } else if ((rc = eph_b200_post_force_begin(h, h->res_x.p, h->res_v.p, dxi, ntimestep, EPH_B200_DEVICE))) return rc;
  h->res_pf_started = false;
  if (h->comm && h->comm_size > 1 && (rc = eph_b200_exchange_ghosts(h))) return rc;
  // the force pass does not wait for the pair forces: it adds into a cleared array of its own, which joins the uploaded
  // forces once both are there (f = f_pair + (f_EPH + f_RNG))
  EPH_CUDA(h, h->res_fe.reserve(3 * std::max<size_t>(nl, 1)));
  if (nl > 0) EPH_CUDA(h, cudaMemsetAsync(h->res_fe.p, 0, 3 * (size_t)nl * sizeof(double), h->stream));
  if ((rc = eph_b200_post_force_end(h, h->res_fe.p, EPH_B200_DEVICE))) return rc;
  if (nl > 0) {
    EPH_CUDA(h, cudaStreamWaitEvent(h->stream, h->f_event, 0));
    add_rows_kernel<<<std::min(blocks_for(3LL * nl, 256), 16 * h->sm_count), 256, 0, h->stream>>>(3LL * nl, h->res_f.p, h->res_fe.p);
    EPH_LAUNCH_CHECK(h);
  }
  h->res_f_valid = true;
  if (nl > 0 && f_out) {
    EPH_CUDA(h, cudaMemcpyAsync(f_out, h->res_f.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

// v_out: HOST [nlocal][3] or NULL (a caller that does not need the velocities on the host this step saves the copy)
int eph_b200_resident_final_integrate(eph_b200_handle *h, const double *mass_by_type, double dtf, double *v_out) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->resident || !h->res_f_valid) return fail(h, EPH_B200_ERR_ARG, "resident_final_integrate: no resident forces (resident_post_force first)");
  if (!mass_by_type) return fail(h, EPH_B200_ERR_ARG, "resident_final_integrate: null masses");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  if (nl == 0) return EPH_B200_OK;
  { const int rc_m = upload_masses(h, mass_by_type); if (rc_m) return rc_m; }
  integrate_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(nl, nullptr, h->res_v.p, h->res_f.p, h->type.p, h->mask.p, h->mass.p, h->cfg.groupbit, 0.0, dtf, 0);
  EPH_LAUNCH_CHECK(h);
  if (v_out) {
    EPH_CUDA(h, cudaMemcpyAsync(v_out, h->res_v.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

// which: 0 x, 1 v, 2 f of the local atoms -> HOST out [nlocal][3]
int eph_b200_resident_get(eph_b200_handle *h, int which, double *out) {
  if (!h || !out) return EPH_B200_ERR_ARG;
  if (!h->resident) return fail(h, EPH_B200_ERR_ARG, "resident_get: resident_upload not called since set_atoms");
  const double *src = which == 0 ? h->res_x.p : which == 1 ? h->res_v.p : which == 2 ? h->res_f.p : nullptr;
  if (!src) return fail(h, EPH_B200_ERR_ARG, "resident_get: bad id %d", which);
  cudaSetDevice(h->cfg.device);
  if (h->nlocal > 0) {
    EPH_CUDA(h, cudaMemcpyAsync(out, src, 3 * (size_t)h->nlocal * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

// end_of_step on the resident velocities
int eph_b200_resident_end_of_step(eph_b200_handle *h, double *E_local) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->resident) return fail(h, EPH_B200_ERR_ARG, "resident_end_of_step: resident_upload not called since set_atoms");
  return eph_b200_end_of_step(h, nullptr, h->res_v.p, E_local, EPH_B200_DEVICE);
}

}  // extern "C"
