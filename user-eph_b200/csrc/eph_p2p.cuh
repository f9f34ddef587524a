// Ghost exchange over NVLink peer memory: the sending kernel gathers the rows other ranks need and stores them straight
// into the receivers' memory, then raises a flag there; the receiving kernel scatters them into the ghost slots.  One
// kernel on each side and no library call in between, for the two per-step exchanges of the path:
//   kind 0  {rho, Wx, Wy, Wz} (+ the injected Gaussians) between the two sweeps   (reference: the RHO / WI / XI forward
//           comms, fix_eph.cpp:743-744, :863-871)
//   kind 1  {x, v} of the ghost atoms ahead of the step                           (LAMMPS' Comm::forward_comm with
//           ghost_velocity, fix_eph.cpp:82; for callers whose atoms live on the device)
// With NCCL each of them is pack kernel -> grouped ncclSend/ncclRecv (about 40 us for 7 peers, latency-bound: 4 MB)
// -> unpack kernel.
//
// Every rank owns one WINDOW (cudaMalloc, exported with cudaIpcGetMemHandle at eph_b200_comm_init and mapped by all ranks
// of the node):
//   header   flags[2 kinds][kP2PMaxRanks] (64-bit epochs): flags[k][s] = e says "sender s has finished writing its rows
//            of exchange number e of kind k"
//   regions  one per sender rank, two halves each (epoch parity); a half holds the kind-0 rows (32 B, then 24 B per row
//            for xi) followed by the kind-1 rows (48 B) of that sender
// Exchange e of a kind goes into half e & 1.  A sender may overwrite a half two exchanges later: by then it has received
// the receiver's exchange e + 1, which the receiver sent (stream order) after it had read exchange e -- every pair of
// peers signals in both directions in every exchange, whether or not it has rows to send.  All ranks call the exchanges
// in lockstep (they are collective), so the epochs agree without being communicated.
#pragma once

#include <cuda_runtime.h>

namespace ephb {

constexpr int kP2PMaxPeers = 32;
constexpr int kP2PMaxRanks = 512;
constexpr size_t kP2PHeaderBytes = 2 * kP2PMaxRanks * sizeof(unsigned long long);
constexpr unsigned kStatusP2PTimeout = 1u << 8;

struct P2PMap {
  int n;                              // peers of this rank
  int my_rank;
  int rank[kP2PMaxPeers];
  int send_off[kP2PMaxPeers + 1];     // rows to peer p: [send_off[p], send_off[p + 1]) of the send index list
  int recv_off[kP2PMaxPeers + 1];     // rows from peer p: the same of the receive slot list
  char *remote[kP2PMaxPeers];         // peer p's window, mapped into this process
  char *local;                        // own window
  size_t region_bytes;
  unsigned long long timeout_ns;      // how long the receiving side polls for a peer's flag before it gives up
};

__host__ __device__ inline size_t p2p_kind1_offset(int rows) { return ((size_t)rows * 56 + 255) / 256 * 256; }
__host__ __device__ inline size_t p2p_half_bytes_needed(int rows) { return p2p_kind1_offset(rows) + (size_t)rows * 48; }

#ifndef EPHA_HOST_EMULATION
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kP2PTimeoutNs = 5000000000ull;      // default; EPH_B200_P2P_TIMEOUT_MS
#else
// host build of the tests: the windows are shared-memory objects of the ranks' processes
inline unsigned long long ld_acquire_sys_u64(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_sys_u64(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned long long global_timer_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
constexpr unsigned long long kP2PTimeoutNs = 120000000000ull;   // default on the (slow) host build
#endif

__device__ __forceinline__ int p2p_peer_of(const int *off, int n, int t) {
  int p = 0;
  while (p + 1 < n && t >= off[p + 1]) ++p;
  return p;
}

// after the rows: every block makes its stores visible system-wide, the last one to finish raises this rank's flag in
// every peer's window
__device__ __forceinline__ void p2p_signal(const P2PMap &m, int kind, unsigned long long epoch, unsigned *done_blocks) {
  __shared__ bool last;
  __syncthreads();   // the block's stores happen before thread 0's fence, which makes them visible system-wide (cumulativity)
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(done_blocks, 1u);
    last = prev == gridDim.x - 1;
    if (last) *done_blocks = 0u;   // every block has counted itself: ready for the next launch
  }
  __syncthreads();
  if (last && threadIdx.x < m.n) {
    __threadfence_system();
    unsigned long long *flag = reinterpret_cast<unsigned long long *>(m.remote[threadIdx.x]) + (size_t)kind * kP2PMaxRanks + m.my_rank;
    st_release_sys_u64(flag, epoch);
  }
}

// the receiving side: the first threads of every block poll the peers' flags (bounded: a peer that never arrives sets a
// status bit instead of hanging the device; after that nobody waits again)
__device__ __forceinline__ void p2p_wait(const P2PMap &m, int kind, unsigned long long epoch, unsigned *status) {
  if (threadIdx.x < m.n && !(*status & kStatusP2PTimeout)) {
    const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(m.local) + (size_t)kind * kP2PMaxRanks + m.rank[threadIdx.x];
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys_u64(flag) < epoch) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > m.timeout_ns) { atomicOr(status, kStatusP2PTimeout); break; }
    }
  }
  __syncthreads();
}

// kind 0: row t of the send list -> {rho, W} (and xi) of that atom into the receiver's half
__global__ void __launch_bounds__(256) p2p_send_payload_kernel(P2PMap m, int ns, const int *__restrict__ index, const double *__restrict__ rho,
                                                               const double4 *__restrict__ W4, const double *__restrict__ xi,
                                                               unsigned long long epoch, unsigned *done_blocks) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += gridDim.x * blockDim.x) {
    const int p = p2p_peer_of(m.send_off, m.n, t);
    const int rows = m.send_off[p + 1] - m.send_off[p], r = t - m.send_off[p];
    char *half = m.remote[p] + kP2PHeaderBytes + (size_t)m.my_rank * m.region_bytes + (size_t)(epoch & 1ull) * (m.region_bytes / 2);
    const int a = index[t];
    const double4 W = W4[a];
    reinterpret_cast<double4 *>(half)[r] = make_double4(rho[a], W.x, W.y, W.z);
    if (xi) {
      double *q = reinterpret_cast<double *>(half + (size_t)rows * 32) + 3 * (size_t)r;
      q[0] = xi[3 * (size_t)a]; q[1] = xi[3 * (size_t)a + 1]; q[2] = xi[3 * (size_t)a + 2];
    }
  }
  p2p_signal(m, 0, epoch, done_blocks);
}

// kind 1: {x, v} of the atoms other ranks hold as ghosts.  A row is three 16-byte pieces {x0 x1 | x2 v0 | v1 v2}; one
// thread per PIECE, so that consecutive threads store consecutive 16 bytes of the receiver's window (a thread per row
// would store 16 bytes at a stride of 48)
__global__ void __launch_bounds__(256) p2p_send_xv_kernel(P2PMap m, int ns, const int *__restrict__ index, const double *__restrict__ x,
                                                          const double *__restrict__ v, unsigned long long epoch, unsigned *done_blocks) {
  const long long pieces = 3LL * ns;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < pieces; c += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(c / 3), part = (int)(c - 3LL * t);
    const int p = p2p_peer_of(m.send_off, m.n, t);
    const int rows = m.send_off[p + 1] - m.send_off[p], r = t - m.send_off[p];
    char *half = m.remote[p] + kP2PHeaderBytes + (size_t)m.my_rank * m.region_bytes + (size_t)(epoch & 1ull) * (m.region_bytes / 2);
    double2 *q = reinterpret_cast<double2 *>(half + p2p_kind1_offset(rows)) + 3 * (size_t)r + part;
    const size_t a = 3 * (size_t)index[t];
    *q = part == 0 ? make_double2(x[a], x[a + 1]) : part == 1 ? make_double2(x[a + 2], v[a]) : make_double2(v[a + 1], v[a + 2]);
  }
  p2p_signal(m, 1, epoch, done_blocks);
}

__global__ void __launch_bounds__(256) p2p_recv_payload_kernel(P2PMap m, int nr, const int *__restrict__ slot, double *__restrict__ rho,
                                                               double4 *__restrict__ W4, double *__restrict__ xi, int ntotal,
                                                               unsigned long long epoch, unsigned *status) {
  p2p_wait(m, 0, epoch, status);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nr; t += gridDim.x * blockDim.x) {
    const int p = p2p_peer_of(m.recv_off, m.n, t);
    const int rows = m.recv_off[p + 1] - m.recv_off[p], r = t - m.recv_off[p];
    const char *half = m.local + kP2PHeaderBytes + (size_t)m.rank[p] * m.region_bytes + (size_t)(epoch & 1ull) * (m.region_bytes / 2);
    const int a = slot[t];
    if (a < 0 || a >= ntotal) continue;
    const double2 *row = reinterpret_cast<const double2 *>(half) + 2 * (size_t)r;
    const double2 b0 = __ldcg(row), b1 = __ldcg(row + 1);
    rho[a] = b0.x;
    W4[a] = make_double4(b0.y, b1.x, b1.y, 0.0);
    if (xi) {
      const double *q = reinterpret_cast<const double *>(half + (size_t)rows * 32) + 3 * (size_t)r;
      xi[3 * (size_t)a] = __ldcg(q); xi[3 * (size_t)a + 1] = __ldcg(q + 1); xi[3 * (size_t)a + 2] = __ldcg(q + 2);
    }
  }
}

// record = 1 (first call after a registration): the caller's ghost coordinates are LAMMPS' own, store the image shifts
__global__ void __launch_bounds__(256) p2p_recv_xv_kernel(P2PMap m, int nr, const int *__restrict__ slot, int nlocal, double *__restrict__ x,
                                                          double *__restrict__ v, double *__restrict__ shift, int record,
                                                          unsigned long long epoch, unsigned *status) {
  p2p_wait(m, 1, epoch, status);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nr; t += gridDim.x * blockDim.x) {
    const int p = p2p_peer_of(m.recv_off, m.n, t);
    const int rows = m.recv_off[p + 1] - m.recv_off[p], r = t - m.recv_off[p];
    const char *half = m.local + kP2PHeaderBytes + (size_t)m.rank[p] * m.region_bytes + (size_t)(epoch & 1ull) * (m.region_bytes / 2);
    const double2 *q = reinterpret_cast<const double2 *>(half + p2p_kind1_offset(rows)) + 3 * (size_t)r;
    const double2 q0 = __ldcg(q), q1 = __ldcg(q + 1), q2 = __ldcg(q + 2);
    const double bx[3] = {q0.x, q0.y, q1.x}, bv[3] = {q1.y, q2.x, q2.y};
    const size_t a = 3 * (size_t)slot[t], c = 3 * (size_t)(slot[t] - nlocal);
  #pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (record) shift[c + d] = x[a + d] - bx[d];
      else { x[a + d] = bx[d] + shift[c + d]; v[a + d] = bv[d]; }
    }
  }
}


}  // namespace ephb
