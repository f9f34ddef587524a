// The reference's two uncorrelated friction models on the device:
//   TTM (eph_model 1, fix_eph.cpp:468-503)  f_EPH_i = -beta(rho_i) v_i
//   PRB (eph_model 2, fix_eph.cpp:505-568)  f_EPH_i = beta(rho_i) (v_i - sum_j rho^{type_j - 1}(r_ij) v_j / rho_i)   for rho_i > 0
// and their common random force f_RNG_i = eta_factor alpha(rho_i) sqrt(T_e(x_i)) xi_i (:490-502, :554-567).
// Both reuse the density pass for rho_i; PRB adds one list sweep for the velocity sum.  PRLCM (3) is not offered: the
// reference indexes its table with `jtype - i` there (fix_eph.cpp:601), i.e. reads out of bounds.
#pragma once

#include "eph_sweeps.cuh"

namespace ephb {

// S_i = sum_j rho^{type_j - 1}(r_ij) v_j over the list entries with r^2 < r_c^2, written to W4.  rho(r) is the
// spline of the file's knots in r (eph_beta.h:157-162), not the rho(r^2) table of the density pass, and -- a
// reference quirk kept for parity -- it is indexed with the LAMMPS type, not with type_map (fix_eph.cpp:530).
template <int LANES>
__global__ void __launch_bounds__(256) prb_sweep_kernel(SweepArgs a, const double2 *__restrict__ rho_r_tab, double inv_dr) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;
  const bool inner = a.walk_mode == 2 || (a.walk_mode == 1 && *a.inner_invalid == 0u);
  const int *__restrict__ list = inner ? a.ineigh : a.neigh;

  for (int i = blockIdx.x * groups_per_block + group_in_block; i < a.nlocal; i += gridDim.x * groups_per_block) {
    const double4 pi = ld256(a.pv + kPvStride * (size_t)i);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    if (double_to_bits(pi.w) & kBitGroup) {
      const RowWalk rw = inner ? walk_tile<LANES>(a, i, lane) : walk_csr<LANES>(a, i, sub);
      const int nn = inner ? a.icount[i] : static_cast<int>(a.offsets[i + 1] - a.offsets[i]);
      const int *__restrict__ lp = list + rw.first;
      int slot = 0;
      for (int k = sub; k < nn; k += LANES, slot += rw.stride) {
        const int j = ld_stream(lp + slot) & kNeighMask;
        const double4 *rec = a.pv + kPvStride * (size_t)j;
        const double4 pj = ld256(rec);
        const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
        const double r2 = ex * ex + ey * ey + ez * ez;
        if (r2 < a.r_cutoff_sq) {
          const double4 vj = ld256(rec + 1);
          const unsigned tj = (double_to_bits(pj.w) >> kTypeShift) & 0xFFu;
          const double rho_j = spline_eval(rho_r_tab + 2 * (size_t)tj * a.n_rho, inv_dr, sqrt(r2));
          sx += rho_j * vj.x; sy += rho_j * vj.y; sz += rho_j * vj.z;
        }
      }
      sx = group_sum<LANES>(sx, gmask); sy = group_sum<LANES>(sy, gmask); sz = group_sum<LANES>(sz, gmask);
    }
    if (sub == 0) a.W4[i] = make_double4(sx, sy, sz, 0.0);
  }
}

struct LegacyArgs {
  int nlocal;
  int model;                             // 1 TTM, 2 PRB
  const double4 *__restrict__ pv;        // {x,y,z,bits | v}
  const double *__restrict__ rho;
  const double4 *__restrict__ S4;        // PRB: velocity sum of prb_sweep_kernel
  const double *__restrict__ xi;         // [nlocal][3]
  const double2 *__restrict__ alpha_tab, *__restrict__ beta_tab;
  int n_beta;
  double inv_drho, rho_cutoff;
  const double *__restrict__ T_e;
  GridGeom grid;
  double eta_factor;
  int do_friction, do_random, add_friction, add_random;
  double *__restrict__ f;                // LAMMPS force array (read-modify-write) or nullptr
  double *__restrict__ f_eph, *__restrict__ f_rng;
};

__global__ void __launch_bounds__(256) legacy_force_kernel(LegacyArgs p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.nlocal) return;
  const double4 pi = p.pv[kPvStride * (size_t)i];
  const unsigned bits = double_to_bits(pi.w);
  double fx = 0, fy = 0, fz = 0, rx = 0, ry = 0, rz = 0;
  if (bits & kBitGroup) {
    const double4 vi = p.pv[kPvStride * (size_t)i + 1];
    const double rho = p.rho[i];
    const size_t tab = 2 * (size_t)(bits & kElemMask) * p.n_beta;
    if (p.do_friction) {
      double beta = 0.0;  // eph_beta.h:171-184
      if (!(rho > p.rho_cutoff)) beta = spline_eval(p.beta_tab + tab, p.inv_drho, rho);
      if (p.model == 1) {            // fix_eph.cpp:478-487
        fx = -beta * vi.x; fy = -beta * vi.y; fz = -beta * vi.z;
      } else if (rho > 0) {          // fix_eph.cpp:520-548
        const double4 S = p.S4[i];
        fx = (vi.x - S.x / rho) * beta; fy = (vi.y - S.y / rho) * beta; fz = (vi.z - S.z / rho) * beta;
      }
    }
    if (p.do_random) {               // fix_eph.cpp:490-502, :554-567
      double alpha = 0.0;            // eph_beta.h:186-198
      if (!(rho > p.rho_cutoff)) alpha = spline_eval(p.alpha_tab + tab, p.inv_drho, rho);
      const double var = p.eta_factor * alpha * sqrt(p.T_e[grid_index(p.grid, pi.x, pi.y, pi.z)]);
      rx = var * p.xi[3 * (size_t)i]; ry = var * p.xi[3 * (size_t)i + 1]; rz = var * p.xi[3 * (size_t)i + 2];
    }
  }
  const size_t o = 3 * (size_t)i;
  if (p.do_friction) { p.f_eph[o] = fx; p.f_eph[o + 1] = fy; p.f_eph[o + 2] = fz; }
  if (p.do_random) { p.f_rng[o] = rx; p.f_rng[o + 1] = ry; p.f_rng[o + 2] = rz; }
  if (p.f != nullptr) {  // f += f_EPH (+ f_RNG) for every local atom (fix_eph.cpp:892-906)
    double ax = 0, ay = 0, az = 0;
    if (p.add_friction) { ax += fx; ay += fy; az += fz; }
    if (p.add_random) { ax += rx; ay += ry; az += rz; }
    p.f[o] += ax; p.f[o + 1] += ay; p.f[o + 2] += az;
  }
}

// `fix eph/coloured/exp` (fix_eph_coloured_exp.cpp): exponential memory kernel on the forces of model 4.  For group atoms
// with rho_i > 0 (the only ones the reference's loops reach, :486, :531, :585)
//   f_dis_i <- f_dis_i (1 - zeta) + zeta f_EPH_i,  f_EPH_i <- f_dis_i     (:563-569, when FRICTION is set)
//   f_sto_i <- f_sto_i (1 - zeta) + zeta f_RNG_i,  f_RNG_i <- f_sto_i     (:619-625, when RANDOM is set)
// then f += f_EPH (+ f_RNG) for every local atom (:664-678).  The force pass ran with its own `f +=` switched off.
struct ColourArgs {
  int nlocal;
  const double4 *__restrict__ pos4;   // {x, y, z, bits}
  const double *__restrict__ rho;
  double zeta;
  int do_friction, do_random, add_friction, add_random;
  double *__restrict__ f_eph, *__restrict__ f_rng, *__restrict__ f_dis, *__restrict__ f_sto;
  double *__restrict__ f;             // LAMMPS force array (read-modify-write) or nullptr
};

__global__ void __launch_bounds__(256) colour_filter_kernel(ColourArgs p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.nlocal) return;
  const bool active = (double_to_bits(p.pos4[i].w) & kBitGroup) && p.rho[i] > 0;
  const size_t o = 3 * (size_t)i;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double fe = p.f_eph[o + d], fr = p.f_rng[o + d];
    if (active && p.do_friction) {
      fe = p.f_dis[o + d] * (1. - p.zeta) + p.zeta * fe;
      p.f_dis[o + d] = fe;
      p.f_eph[o + d] = fe;
    }
    if (active && p.do_random) {
      fr = p.f_sto[o + d] * (1. - p.zeta) + p.zeta * fr;
      p.f_sto[o + d] = fr;
      p.f_rng[o + d] = fr;
    }
    if (p.f != nullptr) {
      double a = 0.0;
      if (p.add_friction) a += fe;
      if (p.add_random) a += fr;
      p.f[o + d] += a;
    }
  }
}

}  // namespace ephb
