// Neighbour-list sweeps of the `fix eph` hot path (model PRL) for sm_100a.
//
// Three passes over the neighbour list, separated by the two ghost broadcasts
// the algorithm needs (rho, then u = alpha/rho * w):
//   rho_sweep       rho_i = sum_j rho^{t_j}(r^2)            (fix_eph.cpp:431-466)
//                   + per-step compaction of the list to the in-cutoff pairs
//   w_rng_sweep     w_i (fix_eph.cpp:702-741) and f_RNG_i (:791-836)
//   friction_sweep  f_EPH_i (fix_eph.cpp:748-787)
// LANES lanes of a warp share one atom; partial sums are combined with
// shuffles.  Algebra: with s = alpha(rho)/rho per atom,
// u = s*w and z = s*xi, every pair term is
//   [ rho^{t_j}(r^2) (e.a_i) - rho^{t_i}(r^2) (e.a_j) ] / r^2 * e ,  a in {u, z}
// so alpha is evaluated once per atom, not once per pair.
//
// The kernels are bound by L1TEX wavefronts (ncu, profiles/r1_*), so the data
// path is organised around them: every per-atom record is 32 bytes and fetched
// with ONE 256-bit load; the rho sweep walks a two-level Verlet list (an inner
// list with a small skin, rebuilt on the device from LAMMPS' list, falling back
// to LAMMPS' list whenever a device-side displacement check invalidates it).
#pragma once

#include "eph_device.cuh"

namespace ephb {

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
  int *__restrict__ ineigh;               // inner list, same row starts
  int *__restrict__ icount;               // [nlocal] inner-list lengths
  const unsigned *__restrict__ inner_invalid;  // device flag: != 0 -> inner list must not be used
  int use_inner;                          // an inner list exists
  int *__restrict__ cneigh;               // compacted in-cutoff list, same row starts
  int *__restrict__ ccount;               // [nlocal] in-cutoff pairs per atom
  const double4 *__restrict__ pos4;       // [ntotal] x,y,z,bits
  const double4 *__restrict__ v4;         // [ntotal] velocity
  const double4 *__restrict__ z4;         // [ntotal] s * xi
  double4 *__restrict__ u4;               // [ntotal] s * w
  const double *__restrict__ s;           // [ntotal] alpha(rho)/rho
  double *__restrict__ rho;               // [ntotal]
  double *__restrict__ w;                 // [nlocal][3]
  double *__restrict__ f_eph;             // [nlocal][3]
  double *__restrict__ f_rng;             // [nlocal][3]
  const double *__restrict__ T_e;         // grid temperatures
  GridGeom grid;
  double eta_factor;
  int do_friction, do_random;
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

// BUILD = this launch also (re)builds the inner list from LAMMPS' list.
template <int LANES, int TAB, bool BUILD>
__global__ void __launch_bounds__(256) rho_sweep_kernel(SweepArgs a) {
  extern __shared__ double2 s_tab[];
  const RhoTable<TAB> tab = stage_tables<TAB>(a, s_tab);
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int gshift = lane & ~(LANES - 1);
  const unsigned below = (1u << sub) - 1u;
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;
  const bool inner = !BUILD && a.use_inner && (*a.inner_invalid == 0u);
  const int *__restrict__ list = inner ? a.ineigh : a.neigh;

  for (int i = blockIdx.x * groups_per_block + group_in_block; i < a.nlocal; i += gridDim.x * groups_per_block) {
    const double4 pi = ld256(a.pos4 + i);
    const unsigned bi = double_to_bits(pi.w);
    double rho = 0.0;
    int count = 0, icnt = 0;
    if (bi & kBitGroup) {  // atoms outside the fix group keep rho = 0 (fix_eph.cpp:442-445)
      const long long beg = a.offsets[i];
      const int nn = inner ? a.icount[i] : static_cast<int>(a.offsets[i + 1] - beg);
      for (int k0 = 0; k0 < nn; k0 += LANES) {
        const int k = k0 + sub;
        bool in = false, in_inner = false;
        int j = 0;
        double r2 = 0.0;
        unsigned bj = 0;
        if (k < nn) {
          j = list[beg + k] & kNeighMask;
          const double4 pj = ld256(a.pos4 + j);
          const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
          r2 = ex * ex + ey * ey + ez * ez;
          bj = double_to_bits(pj.w);
          in = r2 < a.r_cutoff_sq;  // strict '<' as in fix_eph.cpp:457
          if (BUILD) in_inner = r2 < a.r_inner_sq;
        }
        if (BUILD) {
          const unsigned bal = (__ballot_sync(gmask, in_inner) >> gshift) & lanes_bits<LANES>();
          if (in_inner) a.ineigh[beg + icnt + __popc(bal & below)] = j;
          icnt += __popc(bal);
        }
        const unsigned bal = (__ballot_sync(gmask, in) >> gshift) & lanes_bits<LANES>();
        if (in) {
          a.cneigh[beg + count + __popc(bal & below)] = j;
          rho += tab.eval((bj & kElemMask) * a.n_rho, a.inv_dr_sq, r2);
        }
        count += __popc(bal);
      }
      rho = group_sum<LANES>(rho, gmask);
    }
    if (sub == 0) {
      a.rho[i] = rho;
      a.ccount[i] = count;
      if (BUILD) a.icount[i] = icnt;
    }
  }
}

// w_i and f_RNG_i in one pass over the compacted list: both need only rho
// (through s and the validity bit) and share e, r^2, 1/r^2 and the table
// look-ups.  MULTI = more than one element (two look-ups per pair).
template <int LANES, int TAB, bool MULTI>
__global__ void __launch_bounds__(256) w_rng_sweep_kernel(SweepArgs a) {
  extern __shared__ double2 s_tab[];
  const RhoTable<TAB> tab = stage_tables<TAB>(a, s_tab);
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;

  for (int i = blockIdx.x * groups_per_block + group_in_block; i < a.nlocal; i += gridDim.x * groups_per_block) {
    const double4 pi = ld256(a.pos4 + i);
    const unsigned bi = double_to_bits(pi.w);
    double wx = 0, wy = 0, wz = 0, rx = 0, ry = 0, rz = 0;
    // group atoms with rho_i > 0 only (fix_eph.cpp:704-709, :793-798)
    const bool active = (bi & kBitGroup) && (bi & kBitValid);
    if (active) {
      const double4 vi = ld256(a.v4 + i);
      const double4 zi = ld256(a.z4 + i);
      const int off_i = (bi & kElemMask) * a.n_rho;
      const long long beg = a.offsets[i];
      const int nn = a.ccount[i];
      for (int k = sub; k < nn; k += LANES) {
        const int j = a.cneigh[beg + k];
        const double4 pj = ld256(a.pos4 + j);
        const unsigned bj = double_to_bits(pj.w);
        const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
        const double r2 = ex * ex + ey * ey + ez * ez;
        const double rinv = fast_rcp(r2);
        const double rho_j = tab.eval(MULTI ? (bj & kElemMask) * a.n_rho : off_i, a.inv_dr_sq, r2);
        if (a.do_friction) {  // fix_eph.cpp:726-738; no test on rho_j here
          const double4 vj = ld256(a.v4 + j);
          const double d = ex * (vi.x - vj.x) + ey * (vi.y - vj.y) + ez * (vi.z - vj.z);
          const double g = rho_j * rinv * d;
          wx += g * ex; wy += g * ey; wz += g * ez;
        }
        if (a.do_random && (bj & kBitValid)) {  // fix_eph.cpp:811-826
          const double4 zj = ld256(a.z4 + j);
          const double rho_i = MULTI ? tab.eval(off_i, a.inv_dr_sq, r2) : rho_j;
          const double di = ex * zi.x + ey * zi.y + ez * zi.z;
          const double dj = ex * zj.x + ey * zj.y + ez * zj.z;
          const double g = (rho_j * di - rho_i * dj) * rinv;
          rx += g * ex; ry += g * ey; rz += g * ez;
        }
      }
      wx = group_sum<LANES>(wx, gmask); wy = group_sum<LANES>(wy, gmask); wz = group_sum<LANES>(wz, gmask);
      rx = group_sum<LANES>(rx, gmask); ry = group_sum<LANES>(ry, gmask); rz = group_sum<LANES>(rz, gmask);
    }
    if (sub == 0) {
      const double si = active ? a.s[i] : 0.0;
      wx *= si; wy *= si; wz *= si;  // w_i = alpha_i/rho_i * sum (prescaler of fix_eph.cpp:727)
      if (a.do_friction) {
        a.w[3 * (size_t)i + 0] = wx; a.w[3 * (size_t)i + 1] = wy; a.w[3 * (size_t)i + 2] = wz;
        a.u4[i] = make_double4(si * wx, si * wy, si * wz, 0.0);
      }
      if (a.do_random) {
        double var = 0.0;
        if (active) {  // fix_eph.cpp:829-833: nearest-cell T_e, eta_factor = sqrt(2 k_B / dt)
          const double Te = a.T_e[grid_index(a.grid, pi.x, pi.y, pi.z)];
          var = a.eta_factor * sqrt(Te);
        }
        a.f_rng[3 * (size_t)i + 0] = rx * var; a.f_rng[3 * (size_t)i + 1] = ry * var; a.f_rng[3 * (size_t)i + 2] = rz * var;
      }
    }
  }
}

template <int LANES, int TAB, bool MULTI>
__global__ void __launch_bounds__(256) friction_sweep_kernel(SweepArgs a) {
  extern __shared__ double2 s_tab[];
  const RhoTable<TAB> tab = stage_tables<TAB>(a, s_tab);
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;

  for (int i = blockIdx.x * groups_per_block + group_in_block; i < a.nlocal; i += gridDim.x * groups_per_block) {
    const double4 pi = ld256(a.pos4 + i);
    const unsigned bi = double_to_bits(pi.w);
    double fx = 0, fy = 0, fz = 0;
    if ((bi & kBitGroup) && (bi & kBitValid)) {  // fix_eph.cpp:749-754
      const double4 ui = ld256(a.u4 + i);
      const int off_i = (bi & kElemMask) * a.n_rho;
      const long long beg = a.offsets[i];
      const int nn = a.ccount[i];
      for (int k = sub; k < nn; k += LANES) {
        const int j = a.cneigh[beg + k];
        const double4 pj = ld256(a.pos4 + j);
        const unsigned bj = double_to_bits(pj.w);
        if (!(bj & kBitValid)) continue;  // rho_j > 0 required, fix_eph.cpp:768
        const double4 uj = ld256(a.u4 + j);
        const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
        const double r2 = ex * ex + ey * ey + ez * ez;
        const double rinv = fast_rcp(r2);
        const double rho_j = tab.eval(MULTI ? (bj & kElemMask) * a.n_rho : off_i, a.inv_dr_sq, r2);
        const double rho_i = MULTI ? tab.eval(off_i, a.inv_dr_sq, r2) : rho_j;
        const double di = ex * ui.x + ey * ui.y + ez * ui.z;
        const double dj = ex * uj.x + ey * uj.y + ez * uj.z;
        const double g = (rho_j * di - rho_i * dj) * rinv;
        fx -= g * ex; fy -= g * ey; fz -= g * ez;  // friction is negative, fix_eph.cpp:781-784
      }
      fx = group_sum<LANES>(fx, gmask); fy = group_sum<LANES>(fy, gmask); fz = group_sum<LANES>(fz, gmask);
    }
    if (sub == 0) {
      a.f_eph[3 * (size_t)i + 0] = fx; a.f_eph[3 * (size_t)i + 1] = fy; a.f_eph[3 * (size_t)i + 2] = fz;
    }
  }
}

}  // namespace ephb
