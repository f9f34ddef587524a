// Neighbour-list sweeps of the `fix eph` hot path (model PRL) for sm_100a.
//
// The reference walks the full list four times per step (fix_eph.cpp:431-466,
// :702-741, :748-787, :791-836).  Here the work is two passes:
//
//   density_sweep   rho_i = sum_j rho^{t_j}(r^2)                      (fix_eph.cpp:450-461)
//                   W_i   = sum_j g_ij (e.(v_i - v_j)) e              (the sum of :713-739 without its prefactor)
//                   + this step's in-cutoff pair list and the pair weights g_ij = rho^{t_j}(r^2)/r^2
//   force_sweep     f_EPH_i (fix_eph.cpp:758-785) and f_RNG_i (:802-833) from the cached pairs
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
// wavefront per distinct 32-byte sector a warp instruction touches -- so the
// data path is organised around sectors: every per-atom record is one aligned
// sector fetched with ONE 256-bit load; LANES lanes share an atom so that the
// atoms of a warp (spatial neighbours) hit common sectors; the density pass
// walks a two-level Verlet list (inner list with a small skin, rebuilt on the
// device from LAMMPS' list and guarded by a device-side displacement check);
// the spline look-up and the reciprocal are done once per pair per step and
// cached (8 bytes per pair, streamed) instead of being redone by every pass.
#pragma once

#include "eph_device.cuh"

// resident CTAs per SM the sweeps are compiled for (registers: 64 -> 4 CTAs of 256 threads; asking for more spills)
#ifndef EPH_MINB_DENSITY
#define EPH_MINB_DENSITY 4
#endif
#ifndef EPH_MINB_FORCE
#define EPH_MINB_FORCE 4
#endif

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
  int *__restrict__ cneigh;               // this step's in-cutoff pairs, same row starts
  int *__restrict__ ccount;               // [nlocal] in-cutoff pairs per atom
  double *__restrict__ gpair;             // rho^{t_j}(r^2)/r^2 per in-cutoff pair, same indexing as cneigh
  double *__restrict__ gpair_i;           // rho^{t_i}(r^2)/r^2 (only when there is more than one element)
  const double4 *__restrict__ pos4;       // [ntotal] x,y,z,bits
  const double4 *__restrict__ v4;         // [ntotal] velocity
  const double4 *__restrict__ z4;         // [ntotal] s * xi
  const double4 *__restrict__ u4;         // [ntotal] s * w
  double4 *__restrict__ W4;               // [nlocal] sum_j g (e.(v_i - v_j)) e
  double *__restrict__ rho;               // [ntotal]
  double *__restrict__ f;                 // [nlocal][3] LAMMPS force array (read-modify-write) or nullptr
  double *__restrict__ f_eph;             // [nlocal][3]
  double *__restrict__ f_rng;             // [nlocal][3]
  const double *__restrict__ T_e;         // grid temperatures
  GridGeom grid;
  double eta_factor;
  int do_friction, do_random, add_friction, add_random;
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
// MULTI = more than one element: two table look-ups per pair.
template <int LANES, int TAB, bool BUILD, bool MULTI>
__global__ void __launch_bounds__(256, EPH_MINB_DENSITY) density_sweep_kernel(SweepArgs a) {
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
    double rho = 0.0, wx = 0.0, wy = 0.0, wz = 0.0;
    int count = 0, icnt = 0;
    if (bi & kBitGroup) {  // atoms outside the fix group keep rho = 0 (fix_eph.cpp:442-445) and w = 0 (:704)
      const double4 vi = a.do_friction ? ld256(a.v4 + i) : make_double4(0, 0, 0, 0);
      const int off_i = (bi & kElemMask) * a.n_rho;
      const long long beg = a.offsets[i];
      const int nn = inner ? a.icount[i] : static_cast<int>(a.offsets[i + 1] - beg);
      for (int k0 = 0; k0 < nn; k0 += LANES) {
        const int k = k0 + sub;
        bool in = false, in_inner = false;
        int j = 0;
        double r2 = 1.0, ex = 0.0, ey = 0.0, ez = 0.0;
        unsigned bj = 0;
        if (k < nn) {
          j = list[beg + k] & kNeighMask;
          const double4 pj = ld256(a.pos4 + j);
          ex = pj.x - pi.x; ey = pj.y - pi.y; ez = pj.z - pi.z;
          r2 = ex * ex + ey * ey + ez * ez;
          bj = double_to_bits(pj.w);
          in = r2 < a.r_cutoff_sq;  // strict '<' as in fix_eph.cpp:457, :724
          if (BUILD) in_inner = r2 < a.r_inner_sq;
        }
        if (BUILD) {
          const unsigned bal = (__ballot_sync(gmask, in_inner) >> gshift) & lanes_bits<LANES>();
          if (in_inner) a.ineigh[beg + icnt + __popc(bal & below)] = j;
          icnt += __popc(bal);
        }
        const unsigned bal = (__ballot_sync(gmask, in) >> gshift) & lanes_bits<LANES>();
        if (in) {
          const long long slot = beg + count + __popc(bal & below);
          const double rho_j = tab.eval(MULTI ? (bj & kElemMask) * a.n_rho : off_i, a.inv_dr_sq, r2);
          const double rinv = fast_rcp(r2);
          const double g = rho_j * rinv;
          rho += rho_j;
          a.cneigh[slot] = j;
          a.gpair[slot] = g;
          if (MULTI) a.gpair_i[slot] = tab.eval(off_i, a.inv_dr_sq, r2) * rinv;
          if (a.do_friction) {  // fix_eph.cpp:726-738 without the per-atom prefactor alpha_i/rho_i; no test on rho_j
            const double4 vj = ld256(a.v4 + j);
            const double d = g * (ex * (vi.x - vj.x) + ey * (vi.y - vj.y) + ez * (vi.z - vj.z));
            wx += d * ex; wy += d * ey; wz += d * ez;
          }
        }
        count += __popc(bal);
      }
      rho = group_sum<LANES>(rho, gmask);
      if (a.do_friction) {
        wx = group_sum<LANES>(wx, gmask); wy = group_sum<LANES>(wy, gmask); wz = group_sum<LANES>(wz, gmask);
      }
    }
    if (sub == 0) {
      a.rho[i] = rho;
      a.ccount[i] = count;
      a.W4[i] = make_double4(wx, wy, wz, 0.0);
      if (BUILD) a.icount[i] = icnt;
    }
  }
}

// f_EPH_i and f_RNG_i from the cached in-cutoff pairs; no table look-up, no reciprocal, no distance test.
template <int LANES, bool MULTI>
__global__ void __launch_bounds__(256, EPH_MINB_FORCE) force_sweep_kernel(SweepArgs a) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;

  for (int i = blockIdx.x * groups_per_block + group_in_block; i < a.nlocal; i += gridDim.x * groups_per_block) {
    const double4 pi = ld256(a.pos4 + i);
    const unsigned bi = double_to_bits(pi.w);
    double fx = 0, fy = 0, fz = 0, rx = 0, ry = 0, rz = 0;
    // group atoms with rho_i > 0 only (fix_eph.cpp:749-754, :793-798)
    const bool active = (bi & kBitGroup) && (bi & kBitValid);
    if (active) {
      const double4 ui = a.do_friction ? ld256(a.u4 + i) : make_double4(0, 0, 0, 0);
      const double4 zi = a.do_random ? ld256(a.z4 + i) : make_double4(0, 0, 0, 0);
      const long long beg = a.offsets[i];
      const int nn = a.ccount[i];
      for (int k = sub; k < nn; k += LANES) {
        const int j = a.cneigh[beg + k];
        const double gj = a.gpair[beg + k];
        const double gi = MULTI ? a.gpair_i[beg + k] : gj;
        const double4 pj = ld256(a.pos4 + j);
        if (!(double_to_bits(pj.w) & kBitValid)) continue;  // rho_j > 0 required, fix_eph.cpp:768, :811
        const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
        if (a.do_friction) {
          const double4 uj = ld256(a.u4 + j);
          const double di = ex * ui.x + ey * ui.y + ez * ui.z;
          const double dj = ex * uj.x + ey * uj.y + ez * uj.z;
          const double g = gj * di - gi * dj;
          fx -= g * ex; fy -= g * ey; fz -= g * ez;  // friction is negative, fix_eph.cpp:781-784
        }
        if (a.do_random) {
          const double4 zj = ld256(a.z4 + j);
          const double di = ex * zi.x + ey * zi.y + ez * zi.z;
          const double dj = ex * zj.x + ey * zj.y + ez * zj.z;
          const double g = gj * di - gi * dj;
          rx += g * ex; ry += g * ey; rz += g * ez;  // fix_eph.cpp:823-826
        }
      }
      fx = group_sum<LANES>(fx, gmask); fy = group_sum<LANES>(fy, gmask); fz = group_sum<LANES>(fz, gmask);
      rx = group_sum<LANES>(rx, gmask); ry = group_sum<LANES>(ry, gmask); rz = group_sum<LANES>(rz, gmask);
    }
    if (sub == 0) {
      double var = 0.0;
      if (active && a.do_random) {  // fix_eph.cpp:829-833: nearest-cell T_e, eta_factor = sqrt(2 k_B / dt)
        const double Te = a.T_e[grid_index(a.grid, pi.x, pi.y, pi.z)];
        var = a.eta_factor * sqrt(Te);
      }
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
