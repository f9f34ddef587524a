// Neighbour-list sweeps of the `fix eph` hot path (model PRL) for sm_100a.
//
// The reference walks the full list four times per step (fix_eph.cpp:431-466,
// :702-741, :748-787, :791-836).  Here the work is two passes:
//
//   density_sweep   rho_i = sum_j rho^{t_j}(r^2)                      (fix_eph.cpp:450-461)
//                   W_i   = sum_j g_ij (e.(v_i - v_j)) e              (the sum of :713-739 without its prefactor)
//                   + the pair weights g_ij = rho^{t_j}(r^2)/r^2 of every walked list slot (0 beyond the cut-off)
//   force_sweep     f_EPH_i (fix_eph.cpp:758-785) and f_RNG_i (:802-833) from the cached pair weights
//
// Algebra: with the per-atom scalar s = alpha(rho)/rho,
//   w_i = s_i W_i, u = s w, z = s xi, and every pair term of the two forces is
//   [ g_ij (e.a_i) - g_ji (e.a_j) ] e ,  a in {u, z},  g_ji = rho^{t_i}(r^2)/r^2 .
// W_i does not depend on rho, so it rides along with the density pass; w, u, z
// are per-atom products formed between the passes.  One ghost exchange (rho and
// W together) replaces the reference's RHO and WI broadcasts, and alpha is
// evaluated once per atom instead of once per pair.
//
// ncu (profiles/r1_*) shows these kernels bound by L1TEX wavefronts -- about one
// wavefront per distinct 32-byte sector a warp instruction touches -- and, below
// that roof, by the latency of dependent loads.  The data path is organised
// around both:
//   * everything a pass needs about atom j sits in ONE record of consecutive
//     sectors ({pos,bits | v} for the density pass, {pos,bits | u,z} for the
//     force pass), fetched with back-to-back 256-bit loads: one exposed latency
//     per list slot instead of one per array;
//   * the list index (and, in the force pass, the pair weight) of the next
//     slot is fetched one iteration ahead, so the gather never waits for the
//     index stream;
//   * LANES lanes share an atom so that the atoms of a warp (spatial
//     neighbours) hit common sectors;
//   * the density pass walks a two-level Verlet list (inner list with a small
//     skin, rebuilt on the device from LAMMPS' list and guarded by a
//     device-side displacement check);
//   * the inner list and the pair weights are stored warp-tiled: slot t of the
//     32/LANES atoms a warp works on is one run of 32 consecutive entries, so
//     the index and weight streams cost one 128-byte line per warp and
//     iteration instead of one sector per atom;
//   * the spline look-up and the reciprocal are done once per pair per step and
//     cached (8 bytes per walked slot, streamed with evict-first hints).
#pragma once

#include "eph_device.cuh"

// resident CTAs per SM the sweeps are compiled for (registers: 64 -> 4 CTAs of 256 threads; asking for more spills)
#ifndef EPH_MINB_DENSITY
#define EPH_MINB_DENSITY 4
#endif
// force pass: 72 registers hold the working set without spills (spill traffic goes through the same L1 data pipe
// the gathers saturate) -> CTAs of 128 threads, 7 per SM (28 resident warps)
#ifndef EPH_THREADS_FORCE
#define EPH_THREADS_FORCE 128
#endif
#ifndef EPH_MINB_FORCE
#define EPH_MINB_FORCE 7
#endif

namespace ephb {

// Per-atom gather records (32-byte sectors, consecutive in memory).
//   density pass: pv[2a] = {x, y, z, bits}, pv[2a+1] = {vx, vy, vz, 0}
//   force pass:   puz[3a] = {x, y, z, bits}, puz[3a+1] = {ux, uy, uz, zx}, puz[3a+2] = {zy, zz, var, cell}
//                 (var = eta_factor sqrt(T_e(cell of a)) and the cell index, local atoms only, from prep_coupling)
constexpr int kPadIters = 4;   // a freshly built tile is padded to a multiple of this many iterations (eph_packed.cuh)
constexpr int kPvStride = 2;
constexpr int kPuzStride = 3;

struct SweepArgs {
  int nlocal;
  int n_elements;
  int n_rho;
  double inv_dr_sq;
  double r_cutoff_sq;
  double r_inner_sq;                      // (r_c + inner skin)^2
  const double4 *__restrict__ rho_tab4;   // [n_elements][n_rho] {a,b,c,d} (global copy)
  const long long *__restrict__ offsets;  // CSR row starts of LAMMPS' list
  const int *__restrict__ neigh;          // LAMMPS' list (raw entries)
  int *__restrict__ ineigh;               // inner list, warp-tiled (see tile_slot)
  const long long *__restrict__ tile_off; // [ceil(nlocal / (32/LANES)) + 1] first entry of every tile (multiples of 32)
  int *__restrict__ icount;               // [nlocal] inner-list lengths
  const unsigned *__restrict__ inner_invalid;  // device flag: != 0 -> inner list must not be used
  const int *__restrict__ work;           // density pass: tiles to process in this launch (nullptr: all, in order)
  int n_work;                             // density pass: atoms covered by this launch (tiles * 32/LANES, or nlocal)
  int n_boundary;                         // density pass: the first n_boundary work atoms are boundary tiles ...
  unsigned *__restrict__ done_counter;    // ... each finished boundary tile adds 1 here (the exchange waits on it)
  int use_inner;                          // density pass: an inner list exists
  int spec_v;                             // density pass: fetch v_j together with the position when walking the inner list
  int walk_mode;                          // force pass: 0 LAMMPS' list, 1 inner list unless the device flag is set, 2 inner list
  int i_begin, i_end;                     // force pass: atoms [i_begin, i_end) of this launch (i_begin a multiple of 32)
  double *__restrict__ gpair;             // rho^{t_j}(r^2)/r^2 per walked slot (0 beyond r_c), indexed like the walked list
  double *__restrict__ gpair_i;           // rho^{t_i}(r^2)/r^2 (only when there is more than one element)
  const double4 *__restrict__ pv;         // [ntotal][2] density-pass records
  const double4 *__restrict__ puz;        // [ntotal][3] force-pass records
  double4 *__restrict__ W4;               // [nlocal] sum_j g (e.(v_i - v_j)) e
  double *__restrict__ rho;               // [ntotal]
  double *__restrict__ f;                 // [nlocal][3] LAMMPS force array (read-modify-write) or nullptr
  double *__restrict__ f_eph;             // [nlocal][3]
  double *__restrict__ f_rng;             // [nlocal][3]
  int do_friction, do_random, add_friction, add_random;
  int only_fallback;                      // this launch follows a packed sweep (eph_packed.cuh): run only if that one stood
                                          // down because the inner list is invalid
};

// rho(r^2) table access.  TAB = 1: both halves of every record staged in shared
// memory as two bank-friendly double2 arrays; TAB = 0: 256-bit loads through L1.
template <int TAB>
struct RhoTable {
  const double4 *g;
  const double2 *ab, *cd;
  __device__ __forceinline__ double eval(int elem_off, double inv_dx, double x) const {
    const unsigned idx = static_cast<unsigned>(x * inv_dx) + elem_off;  // truncating index, eph_spline.h:138
    if (TAB == 1) {
      const double2 p = ab[idx], q = cd[idx];
      return fma(x, fma(x, fma(x, q.y, q.x), p.y), p.x);
    }
    const double4 c = ld256(g + idx);
    return fma(x, fma(x, fma(x, c.w, c.z), c.y), c.x);
  }
};

template <int TAB>
__device__ __forceinline__ RhoTable<TAB> stage_tables(const SweepArgs &a, double2 *smem) {
  RhoTable<TAB> t;
  t.g = a.rho_tab4;
  t.ab = t.cd = nullptr;
  if (TAB == 1) {
    const int n = a.n_elements * a.n_rho;
    double2 *ab = smem, *cd = smem + n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const double4 c = a.rho_tab4[k];
      ab[k] = make_double2(c.x, c.y);
      cd[k] = make_double2(c.z, c.w);
    }
    __syncthreads();
    t.ab = ab;
    t.cd = cd;
  }
  return t;
}

// Where the slots of one atom live.  LAMMPS' list: CSR row, the group's LANES lanes read LANES consecutive entries
// per iteration.  Inner list: the tile of the 32/LANES atoms a warp works on; iteration t of the whole warp is the
// run [tile_off + 32 t, tile_off + 32 t + 32), lane l reads entry l of it.
struct RowWalk {
  long long first;  // entry of (iteration 0, this lane)
  int stride;       // entries between iterations
};
template <int LANES>
__device__ __forceinline__ RowWalk walk_csr(const SweepArgs &a, int i, int sub) {
  return RowWalk{a.offsets[i] + sub, LANES};
}
template <int LANES>
__device__ __forceinline__ RowWalk walk_tile(const SweepArgs &a, int i, int lane) {
  return RowWalk{a.tile_off[i / (32 / LANES)] + lane, 32};
}

// BUILD = this launch also (re)builds the inner list from LAMMPS' list.
// MULTI = more than one element: two table look-ups per pair.
template <int LANES, int TAB, bool BUILD, bool MULTI>
__global__ void __launch_bounds__(256, EPH_MINB_DENSITY) density_sweep_kernel(SweepArgs a) {
  extern __shared__ double2 s_tab[];
  if (a.only_fallback && *a.inner_invalid == 0u) return;
  const RhoTable<TAB> tab = stage_tables<TAB>(a, s_tab);
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int gshift = lane & ~(LANES - 1);
  const unsigned below = (1u << sub) - 1u;
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;
  const bool inner = !BUILD && a.use_inner && (*a.inner_invalid == 0u);
  const bool spec = inner && a.spec_v;
  const int *__restrict__ list = inner ? a.ineigh : a.neigh;

  // whole warps run the loop (the last tile may be partial): the padding of a freshly built tile needs a warp-wide maximum
  const int n_work_pad = (a.n_work + 32 / LANES - 1) / (32 / LANES) * (32 / LANES);
  for (int w = blockIdx.x * groups_per_block + group_in_block; w < n_work_pad; w += gridDim.x * groups_per_block) {
    // a work list names whole tiles (the warp's 32/LANES atoms), so the lane <-> tile-slot mapping is unchanged
    const int i = a.work ? a.work[w / (32 / LANES)] * (32 / LANES) + (w & (32 / LANES - 1)) : w;
    const bool real = w < a.n_work && i < a.nlocal;
    // the whole header of the atom is requested at once, not behind the record that says whether it is in the group
    double4 pi = make_double4(0, 0, 0, 0), vi = pi;
    RowWalk rw{0, 0};
    int nn = 0;
    if (real) {
      pi = ld256(a.pv + kPvStride * (size_t)i);
      if (a.do_friction) vi = ld256(a.pv + kPvStride * (size_t)i + 1);
      rw = inner ? walk_tile<LANES>(a, i, lane) : walk_csr<LANES>(a, i, sub);
      nn = inner ? a.icount[i] : static_cast<int>(a.offsets[i + 1] - a.offsets[i]);
    }
    const unsigned bi = double_to_bits(pi.w);
    double rho = 0.0, wx = 0.0, wy = 0.0, wz = 0.0;
    int icnt = 0;
    if (real && (bi & kBitGroup)) {  // atoms outside the fix group keep rho = 0 (fix_eph.cpp:442-445) and w = 0 (:704)
      const int off_i = (bi & kElemMask) * a.n_rho;
      const long long tile0 = BUILD ? a.tile_off[i / (32 / LANES)] + gshift : 0;   // this atom's lanes of tile iteration 0
      const int *__restrict__ lp = list + rw.first;
      double *__restrict__ gp = a.gpair + rw.first;
      double *__restrict__ gip = MULTI ? a.gpair_i + rw.first : nullptr;
      int slot = 0;
      int jn = sub < nn ? ld_stream(lp) : 0;   // the index stream runs one slot ahead of the gathers
#pragma unroll 2
      for (int k0 = 0; k0 < nn; k0 += LANES, slot += rw.stride) {
        const int k = k0 + sub;
        const bool have = k < nn;
        const int j = jn & kNeighMask;
        if (k + LANES < nn) jn = ld_stream(lp + slot + rw.stride);
        bool in = false, in_inner = false;
        double g = 0.0, gi = 0.0;
        if (have) {
          const double4 *rec = a.pv + kPvStride * (size_t)j;
          const double4 pj = ld256(rec);
          double4 vj = make_double4(0, 0, 0, 0);
          // inner list: 5 of 6 slots are inside the cut-off, so v_j is fetched together with the position
          // (one latency per slot); LAMMPS' list (3 of 8 inside): only after the distance test
          if (spec && a.do_friction) vj = ld256(rec + 1);
          const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
          const double r2 = ex * ex + ey * ey + ez * ez;
          in = r2 < a.r_cutoff_sq;  // strict '<' as in fix_eph.cpp:457, :724
          if (BUILD) in_inner = r2 < a.r_inner_sq;
          if (in) {
            const unsigned bj = double_to_bits(pj.w);
            const double rho_j = tab.eval(MULTI ? (bj & kElemMask) * a.n_rho : off_i, a.inv_dr_sq, r2);
            const double rinv = fast_rcp(r2);
            g = rho_j * rinv;
            rho += rho_j;
            // the second look-up is only needed when the two atoms are different elements
            if (MULTI) gi = ((bj & kElemMask) * a.n_rho == off_i) ? g : tab.eval(off_i, a.inv_dr_sq, r2) * rinv;
            if (a.do_friction) {  // fix_eph.cpp:726-738 without the per-atom prefactor alpha_i/rho_i; no test on rho_j
              if (!spec) vj = ld256(rec + 1);
              const double d = g * (ex * (vi.x - vj.x) + ey * (vi.y - vj.y) + ez * (vi.z - vj.z));
              wx += d * ex; wy += d * ey; wz += d * ez;
            }
          }
        }
        if (BUILD) {
          const unsigned bal = (__ballot_sync(gmask, in_inner) >> gshift) & lanes_bits<LANES>();
          if (in_inner) {
            const int c = icnt + __popc(bal & below);   // position in this atom's inner list
            const long long dst = tile0 + (long long)(c / LANES) * 32 + (c & (LANES - 1));
            a.ineigh[dst] = j;
            st_stream(a.gpair + dst, g);
            if (MULTI) st_stream(a.gpair_i + dst, gi);
          }
          icnt += __popc(bal);
        } else if (have) {
          st_stream(gp + slot, g);
          if (MULTI) st_stream(gip + slot, gi);
        }
      }
      rho = group_sum<LANES>(rho, gmask);
      if (a.do_friction) {
        wx = group_sum<LANES>(wx, gmask); wy = group_sum<LANES>(wy, gmask); wz = group_sum<LANES>(wz, gmask);
      }
    }
    if (real && sub == 0) {
      a.rho[i] = rho;
      a.W4[i] = make_double4(wx, wy, wz, 0.0);
      if (BUILD) a.icount[i] = icnt;
    }
    if (BUILD) {
      // Padding for the packed sweeps (eph_packed.cuh), which run ONE trip count per warp and prefetch without bounds
      // tests: every atom's slots up to the longest list of the tile, rounded up to kPadIters iterations, hold the atom's own
      // index and a zero pair weight.
      __syncwarp();
      int tmax = icnt;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tmax = max(tmax, __shfl_xor_sync(0xFFFFFFFFu, tmax, o));
      const int padded = (tmax + kPadIters * LANES - 1) / (kPadIters * LANES) * (kPadIters * LANES);
      const long long t0 = a.tile_off[(a.work ? i : w) / (32 / LANES)] + gshift;
      for (int c = icnt + sub; c < padded; c += LANES) {
        const long long dst = t0 + (long long)(c / LANES) * 32 + (c & (LANES - 1));
        a.ineigh[dst] = real ? i : 0;
        st_stream(a.gpair + dst, 0.0);
        if (MULTI) st_stream(a.gpair_i + dst, 0.0);
      }
    }
    // boundary tiles come first in the work list; when the last of them is done the ghost exchange may start
    // (the communication stream waits for the counter with a stream memory operation)
    if (a.done_counter != nullptr && w < a.n_boundary) {   // warp-uniform: a warp owns exactly one tile
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(a.done_counter, 1u);
      }
    }
  }
}

// f_EPH_i and f_RNG_i from the cached pair weights; no table look-up, no reciprocal, no distance test.
template <int LANES, bool MULTI>
__global__ void __launch_bounds__(EPH_THREADS_FORCE, EPH_MINB_FORCE) force_sweep_kernel(SweepArgs a) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;
  if (a.only_fallback && *a.inner_invalid == 0u) return;
  // the list whose slots the density pass of this step filled
  const bool inner = a.walk_mode == 2 || (a.walk_mode == 1 && *a.inner_invalid == 0u);
  const int *__restrict__ list = inner ? a.ineigh : a.neigh;

  for (int i = a.i_begin + blockIdx.x * groups_per_block + group_in_block; i < a.i_end; i += gridDim.x * groups_per_block) {
    // the whole header of the atom is requested at once (the pass is bound by the latency of dependent loads: the
    // row description must not wait for the record that says whether the atom takes part)
    const double4 *ri = a.puz + kPuzStride * (size_t)i;
    const double4 pi = ld256(ri);
    const double4 qi = ld256(ri + 1);
    const double2 si = ld128(ri + 2);
    const RowWalk rw = inner ? walk_tile<LANES>(a, i, lane) : walk_csr<LANES>(a, i, sub);
    const int nn = inner ? a.icount[i] : static_cast<int>(a.offsets[i + 1] - a.offsets[i]);
    const unsigned bi = double_to_bits(pi.w);
    double fx = 0, fy = 0, fz = 0, rx = 0, ry = 0, rz = 0;
    // group atoms with rho_i > 0 only (fix_eph.cpp:749-754, :793-798)
    const bool active = (bi & kBitGroup) && (bi & kBitValid);
    if (active) {
      const double uix = qi.x, uiy = qi.y, uiz = qi.z, zix = qi.w, ziy = si.x, ziz = si.y;
      const int *__restrict__ lp = list + rw.first;
      const double *__restrict__ gp = a.gpair + rw.first;
      const double *__restrict__ gip = MULTI ? a.gpair_i + rw.first : nullptr;
      int slot = 0;
      int jn = 0;
      double gjn = 0.0, gin = 0.0;
      if (sub < nn) {
        jn = ld_stream(lp);
        gjn = ld_stream(gp);
        if (MULTI) gin = ld_stream(gip);
      }
#pragma unroll 1
      for (int k = sub; k < nn; k += LANES, slot += rw.stride) {
        const int j = jn & kNeighMask;
        const double gj = gjn;
        const double gi = MULTI ? gin : gjn;
        if (k + LANES < nn) {
          jn = ld_stream(lp + slot + rw.stride);
          gjn = ld_stream(gp + slot + rw.stride);
          if (MULTI) gin = ld_stream(gip + slot + rw.stride);
        }
        if (gj == 0.0 && gi == 0.0) continue;  // beyond the cut-off (fix_eph.cpp:768, :811) or a vanishing pair weight
        const double4 *rj = a.puz + kPuzStride * (size_t)j;
        const double4 pj = ld256(rj), qj = ld256(rj + 1);
        double2 sj = make_double2(0, 0);
        if (a.do_random) sj = ld128(rj + 2);
        if (!(double_to_bits(pj.w) & kBitValid)) continue;  // rho_j > 0 required, fix_eph.cpp:768, :811
        const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
        if (a.do_friction) {
          const double di = ex * uix + ey * uiy + ez * uiz;
          const double dj = ex * qj.x + ey * qj.y + ez * qj.z;
          const double g = gj * di - gi * dj;
          fx -= g * ex; fy -= g * ey; fz -= g * ez;  // friction is negative, fix_eph.cpp:781-784
        }
        if (a.do_random) {
          const double di = ex * zix + ey * ziy + ez * ziz;
          const double dj = ex * qj.w + ey * sj.x + ez * sj.y;
          const double g = gj * di - gi * dj;
          rx += g * ex; ry += g * ey; rz += g * ez;  // fix_eph.cpp:823-826
        }
      }
      fx = group_sum<LANES>(fx, gmask); fy = group_sum<LANES>(fy, gmask); fz = group_sum<LANES>(fz, gmask);
      rx = group_sum<LANES>(rx, gmask); ry = group_sum<LANES>(ry, gmask); rz = group_sum<LANES>(rz, gmask);
    }
    if (sub == 0) {
      double var = 0.0;
      // fix_eph.cpp:829-833: eta_factor sqrt(T_e(nearest cell)), worked out once per atom by prep_coupling
      if (active && a.do_random) var = __ldg(reinterpret_cast<const double *>(ri + 2) + 2);
      rx *= var; ry *= var; rz *= var;
      const size_t o = 3 * (size_t)i;
      if (a.do_friction) { a.f_eph[o] = fx; a.f_eph[o + 1] = fy; a.f_eph[o + 2] = fz; }
      if (a.do_random) { a.f_rng[o] = rx; a.f_rng[o + 1] = ry; a.f_rng[o + 2] = rz; }
      // f += f_EPH (+ f_RNG) for every local atom, grouped or not (fix_eph.cpp:892-906)
      if (a.f != nullptr) {
        double ax = 0, ay = 0, az = 0;
        if (a.add_friction) { ax += fx; ay += fy; az += fz; }
        if (a.add_random) { ax += rx; ay += ry; az += rz; }
        a.f[o] += ax; a.f[o + 1] += ay; a.f[o + 2] += az;
      }
    }
  }
}

}  // namespace ephb
