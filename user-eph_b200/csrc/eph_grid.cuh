// Electronic-temperature grid kernels: EPH_FDM::solve (reference
// eph_fdm.h:267-400) as a fused, double-buffered explicit stencil.
#pragma once

#include "eph_device.cuh"

namespace ephb {

struct GridArgs {
  int nx, ny, nz;
  long long ncell;
  const double *__restrict__ T_in;
  double *__restrict__ T_out;
  double *__restrict__ dT_e;
  const double *__restrict__ S_e;
  const double *__restrict__ rho_e;
  const double *__restrict__ C_e;
  const double *__restrict__ kappa_e;
  const short *__restrict__ flag;
  const unsigned short *__restrict__ t_dyn;
  // temperature dependent tables (may be null)
  const double *__restrict__ E_e_T;
  int n_T;
  double dT;
  double inv_dx2, inv_dy2, inv_dz2;  // 1/dx^2 ...
  double inner_dt;
  int clear_source;  // last sub-step: zero dT_e (sync_after, eph_fdm.h:484-487)
  unsigned *__restrict__ status;
  // planes [z_begin, z_end) are updated by this launch (the whole grid, or one rank's slab of a sharded solve:
  // neighbours are still read with the periodic wrap of the full grid, the halo planes being current in T_in)
  int z_begin, z_end;
};

// EPH_Linear::operator() and reverse_lookup (reference eph_linear.h:40-62)
__device__ __forceinline__ double linear_eval(const double *__restrict__ y, int n, double dx, double x) {
  size_t idx = static_cast<size_t>(x / dx);
  if (idx < static_cast<size_t>(n)) {
    double dy = (idx + 1 < static_cast<size_t>(n)) ? (y[idx + 1] - y[idx]) / dx : 0.0;
    return y[idx] + dy * (x - idx * dx);
  }
  return 0.;
}
__device__ __forceinline__ double linear_reverse(const double *__restrict__ y, int n, double dx, double yv) {
  int lo = 0, hi = n;  // upper_bound: first element > yv
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (!(yv < y[mid])) lo = mid + 1; else hi = mid;
  }
  if (lo != n && lo > 0) {
    int idx = lo - 1;
    double dy = (y[idx + 1] - y[idx]) / dx;
    return idx * dx + 1. / dy * (yv - y[idx]);
  }
  return 0.;
}

// One explicit sub-step for every cell: 7-point variable-kappa stencil with
// periodic wrap and zero-derivative wall substitution (eph_fdm.h:319-367), the
// update of DYNAMIC cells (:371-388) and the clamp at zero (:391-394), fused so
// each field is read once per sub-step.
__global__ void __launch_bounds__(256) fdm_substep_kernel(GridArgs g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = g.z_begin + blockIdx.z * blockDim.z + threadIdx.z;
  if (i >= g.nx || j >= g.ny || k >= g.z_end) return;
  const long long sx = 1, sy = g.nx, sz = (long long)g.nx * g.ny;
  const long long r = i + j * sy + k * sz;
  const short fr = g.flag[r];
  double T = g.T_in[r];
  if (fr == 1) {  // only DYNAMIC cells change; walls (2) and constant cells (0) keep T
    const double kr = g.kappa_e[r];
    double ddT = 0.0;
    {
      long long p = (i > 0) ? r - sx : r + (g.nx - 1) * sx;
      long long q = (i < g.nx - 1) ? r + sx : r - (g.nx - 1) * sx;
      if (g.flag[q] == 2) q = r; else if (g.flag[p] == 2) p = r;
      const double Tq = g.T_in[q], Tp = g.T_in[p];
      ddT += (g.kappa_e[q] - g.kappa_e[p]) * (Tq - Tp) * g.inv_dx2 * 0.25;
      ddT += kr * ((Tq + Tp - 2.0 * T) * g.inv_dx2);
    }
    {
      long long p = (j > 0) ? r - sy : r + (g.ny - 1) * sy;
      long long q = (j < g.ny - 1) ? r + sy : r - (g.ny - 1) * sy;
      if (g.flag[q] == 2) q = r; else if (g.flag[p] == 2) p = r;
      const double Tq = g.T_in[q], Tp = g.T_in[p];
      ddT += (g.kappa_e[q] - g.kappa_e[p]) * (Tq - Tp) * g.inv_dy2 * 0.25;
      ddT += kr * ((Tq + Tp - 2.0 * T) * g.inv_dy2);
    }
    {
      long long p = (k > 0) ? r - sz : r + (g.nz - 1) * sz;
      long long q = (k < g.nz - 1) ? r + sz : r - (g.nz - 1) * sz;
      if (g.flag[q] == 2) q = r; else if (g.flag[p] == 2) p = r;
      const double Tq = g.T_in[q], Tp = g.T_in[p];
      ddT += (g.kappa_e[q] - g.kappa_e[p]) * (Tq - Tp) * g.inv_dz2 * 0.25;
      ddT += kr * ((Tq + Tp - 2.0 * T) * g.inv_dz2);
    }
    const double src = ddT + g.dT_e[r] + g.S_e[r];
    const double rho = g.rho_e[r];
    if (g.t_dyn[r] == 1 && g.E_e_T != nullptr) {
      double E = linear_eval(g.E_e_T, g.n_T, g.dT, T);
      E += src / rho * g.inner_dt;
      T = linear_reverse(g.E_e_T, g.n_T, g.dT, E);
    } else {
      T += src / (rho * g.C_e[r]) * g.inner_dt;
    }
  }
  if (T < 0.0) {
    T = 0.0;
    atomicOr(g.status, 2u);
  }
  g.T_out[r] = T;
  if (g.clear_source) g.dT_e[r] = 0.0;
}

// C_e(T), kappa_e(T) refresh of temperature-dependent cells (eph_fdm.h:280-287)
__global__ void fdm_refresh_kernel(long long ncell, const double *__restrict__ T, const unsigned short *__restrict__ t_dyn,
                                   const double2 *__restrict__ C_tab, const double2 *__restrict__ K_tab, double inv_dT,
                                   double *__restrict__ C_e, double *__restrict__ kappa_e) {
  long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= ncell) return;
  if (t_dyn[r]) {
    C_e[r] = spline_eval(C_tab, inv_dT, T[r]);
    kappa_e[r] = spline_eval(K_tab, inv_dT, T[r]);
  }
}

// min C_e, min rho_e, max kappa_e over non-constant cells, seeded from cell 0
// unconditionally (eph_fdm.h:290-300).  out = {c_min, rho_min, kappa_max},
// pre-seeded by the host with cell 0's values; combined with ordered-int atomics
// on the raw bits (all three fields are positive).
__global__ void fdm_minmax_kernel(long long ncell, const double *__restrict__ C_e, const double *__restrict__ rho_e,
                                  const double *__restrict__ kappa_e, const short *__restrict__ flag,
                                  unsigned long long *__restrict__ out) {
  double cmin = __longlong_as_double(out[0]), rmin = __longlong_as_double(out[1]), kmax = __longlong_as_double(out[2]);
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < ncell; r += (long long)gridDim.x * blockDim.x) {
    if (r == 0 || flag[r] != 0) {
      cmin = fmin(cmin, C_e[r]);
      rmin = fmin(rmin, rho_e[r]);
      kmax = fmax(kmax, kappa_e[r]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cmin = fmin(cmin, __shfl_xor_sync(0xFFFFFFFFu, cmin, o));
    rmin = fmin(rmin, __shfl_xor_sync(0xFFFFFFFFu, rmin, o));
    kmax = fmax(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&out[0], (unsigned long long)__double_as_longlong(cmin));
    atomicMin(&out[1], (unsigned long long)__double_as_longlong(rmin));
    atomicMax(&out[2], (unsigned long long)__double_as_longlong(kmax));
  }
}

// sum of a field (EPH_FDM::get_T_total, eph_fdm.h:189-196)
__global__ void fdm_sum_kernel(long long ncell, const double *__restrict__ T, double *__restrict__ out) {
  double acc = 0.0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < ncell; r += (long long)gridDim.x * blockDim.x)
    acc += T[r];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

}  // namespace ephb
