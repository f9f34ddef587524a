// TEST INFRASTRUCTURE -- never part of the product, never loaded by it.
//
// A lock-step SIMT stand-in for running the product's CUDA kernels on the host (see cuda_runtime.h in this directory).
// Every CUDA thread of a block is a user-level fiber (ucontext); the blocks of a launch run one after the other.  A
// fiber runs until it reaches a collective (__syncthreads, __shfl*_sync, __ballot_sync, __syncwarp), parks there and the
// scheduler resumes the next one; a collective completes when every participating thread has arrived, exactly as the
// hardware requires, so mismatched masks or divergent barriers show up as a reported dead-lock instead of passing
// silently.  Single host thread: atomics are plain read-modify-writes, results are deterministic.
#pragma once
#include <sched.h>

#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_simt { unsigned x = 0, y = 0, z = 0; };

namespace simt {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  char *stack = nullptr;
  bool done = true;
};

struct Coll {          // one rendezvous point: (warp, mask)
  unsigned mask = 0;
  int count = 0;
  unsigned long long gen = 0;
  uint64_t slots[32];
  uint64_t snap[32];
};

struct State {
  ucontext_t sched;
  Fiber fibers[kMaxThreads];
  int cur = 0, nthreads = 0;
  long long progress = 0;
  std::function<void()> body;
  dim3 block_dim, grid_dim;
  uint3_simt thread_idx, block_idx;
  // __syncthreads
  int active = 0, waiting = 0;
  unsigned long long bar_gen = 0;
  // warp collectives: per warp a short list of masks in use
  std::vector<Coll> colls[kMaxThreads / 32];
  std::vector<unsigned char> dyn_smem;
};

inline State &S() {
  static State s;
  return s;
}

inline void set_thread(int t) {
  State &s = S();
  s.cur = t;
  s.thread_idx.x = t % s.block_dim.x;
  s.thread_idx.y = (t / s.block_dim.x) % s.block_dim.y;
  s.thread_idx.z = t / (s.block_dim.x * s.block_dim.y);
}

inline void yield() {
  State &s = S();
  const int me = s.cur;
  swapcontext(&s.fibers[me].ctx, &s.sched);
  set_thread(me);   // the scheduler ran other fibers in between
}

inline void release_barrier_if_complete() {
  State &s = S();
  if (s.waiting > 0 && s.waiting == s.active) {
    s.waiting = 0;
    ++s.bar_gen;
    ++s.progress;
  }
}

inline void trampoline() {
  State &s = S();
  s.body();
  const int me = s.cur;
  s.fibers[me].done = true;
  --s.active;                      // an exited thread no longer counts for __syncthreads
  ++s.progress;
  release_barrier_if_complete();
  swapcontext(&s.fibers[me].ctx, &s.sched);
}

inline void syncthreads() {
  State &s = S();
  const unsigned long long my = s.bar_gen;
  ++s.waiting;
  ++s.progress;
  release_barrier_if_complete();
  while (s.bar_gen == my) yield();
}

// all lanes named in `mask` (of the calling thread's warp) meet here; returns the values they brought, by lane
inline const uint64_t *collective(unsigned mask, uint64_t v) {
  State &s = S();
  const int lane = s.cur & 31, warp = s.cur >> 5;
  if (!(mask & (1u << lane))) {
    std::fprintf(stderr, "simt: lane %d calls a collective whose mask %08x does not name it\n", lane, mask);
    std::abort();
  }
  // lanes beyond the end of a partial last warp do not exist
  const int lanes_here = s.nthreads - warp * 32 >= 32 ? 32 : s.nthreads - warp * 32;
  const unsigned live = lanes_here >= 32 ? 0xFFFFFFFFu : ((1u << lanes_here) - 1u);
  const unsigned eff = mask & live;
  std::vector<Coll> &list = s.colls[warp];
  size_t k = 0;
  for (; k < list.size(); ++k)
    if (list[k].mask == eff) break;
  if (k == list.size()) {
    list.emplace_back();
    list.back().mask = eff;
  }
  // (the vector may reallocate while this fiber is parked: always re-index, never keep a reference across a yield)
  list[k].slots[lane] = v;
  ++list[k].count;
  ++s.progress;
  const unsigned long long my = list[k].gen;
  if (list[k].count == __builtin_popcount(eff)) {
    std::memcpy(list[k].snap, list[k].slots, sizeof list[k].snap);
    list[k].count = 0;
    ++list[k].gen;
  } else {
    while (s.colls[warp][k].gen == my) yield();
  }
  return s.colls[warp][k].snap;
}

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F &&body) {
  State &s = S();
  const int n = (int)(block.x * block.y * block.z);
  if (n < 1 || n > kMaxThreads) {
    std::fprintf(stderr, "simt: bad block size %d\n", n);
    std::abort();
  }
  s.block_dim = block;
  s.grid_dim = grid;
  s.nthreads = n;
  s.body = body;
  s.dyn_smem.assign(smem_bytes + 128, 0);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.block_idx.x = bx; s.block_idx.y = by; s.block_idx.z = bz;
        s.active = n;
        s.waiting = 0;
        for (auto &l : s.colls) l.clear();
        for (int t = 0; t < n; ++t) {
          Fiber &f = s.fibers[t];
          if (!f.stack) f.stack = static_cast<char *>(std::malloc(kStackBytes));
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, trampoline, 0);
          f.done = false;
        }
        // SIMT_ORDER=reverse | random: the order in which parked fibers are resumed.  Results must not depend on it
        // (beyond the rounding of atomic sums): a kernel that misses a barrier between a shared-memory or global write
        // and another thread's read gives different answers under different orders -- the host build's racecheck.
        static const int order_mode = [] {
          const char *e = std::getenv("SIMT_ORDER");
          return !e ? 0 : std::strcmp(e, "reverse") == 0 ? 1 : std::strcmp(e, "random") == 0 ? 2 : 0;
        }();
        static unsigned long long lcg = 0x9E3779B97F4A7C15ULL;
        int remaining = n;
        while (remaining > 0) {
          const long long before = s.progress;
          remaining = 0;
          int start = 0, step = 1;
          if (order_mode == 1) { start = n - 1; step = -1; }
          if (order_mode == 2) {
            lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
            start = (int)((lcg >> 33) % (unsigned long long)n);
            step = ((lcg >> 20) & 1) ? 1 : -1;
          }
          for (int k = 0, t = start; k < n; ++k, t = (t + step + n) % n) {
            if (s.fibers[t].done) continue;
            set_thread(t);
            swapcontext(&s.sched, &s.fibers[t].ctx);
            if (!s.fibers[t].done) ++remaining;
          }
          if (remaining > 0 && s.progress == before) {
            std::fprintf(stderr, "simt: dead-lock in block (%u,%u,%u): %d thread(s) parked at a collective that cannot complete "
                                 "(mismatched mask or divergent barrier)\n", bx, by, bz, remaining);
            std::abort();
          }
        }
      }
  s.body = nullptr;
}

inline void *dyn_smem() {
  State &s = S();
  return reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(s.dyn_smem.data()) + 127) & ~uintptr_t(127));
}

template <class T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "collective payload wider than 8 bytes");
  uint64_t u = 0;
  std::memcpy(&u, &v, sizeof(T));
  return u;
}
template <class T>
inline T from_bits(uint64_t u) {
  T v;
  std::memcpy(&v, &u, sizeof(T));
  return v;
}

}  // namespace simt

#define threadIdx (simt::S().thread_idx)
#define blockIdx (simt::S().block_idx)
#define blockDim (simt::S().block_dim)
#define gridDim (simt::S().grid_dim)

inline void __syncthreads() { simt::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { simt::collective(mask, 0); }
inline void __threadfence() {}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }   // other PROCESSES read the peer-memory windows
inline void __nanosleep(unsigned) { sched_yield(); }
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
  const int lane = simt::S().cur & 31;
  return simt::from_bits<T>(simt::collective(mask, simt::to_bits(v))[(lane ^ lane_mask) & 31]);
}
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src_lane) {
  return simt::from_bits<T>(simt::collective(mask, simt::to_bits(v))[src_lane & 31]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  const uint64_t *vals = simt::collective(mask, pred ? 1 : 0);
  const int warp = simt::S().cur >> 5;
  const int lanes_here = simt::S().nthreads - warp * 32 >= 32 ? 32 : simt::S().nthreads - warp * 32;
  unsigned r = 0;
  for (int l = 0; l < lanes_here; ++l)
    if ((mask & (1u << l)) && vals[l]) r |= 1u << l;
  return r;
}
