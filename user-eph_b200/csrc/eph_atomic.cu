// C ABI of the B200-native `fix eph/atomic` path (include/eph_b200_atomic.h; reference fix_eph_atomic.cpp, eph_kappa.h).
// A self-contained engine next to the `fix eph` one (eph_b200.cu): same device building blocks (eph_device.cuh), its own
// handle.  LANES lanes of a warp share one atom and walk its row of LAMMPS' full list (CSR); everything a pair needs
// about atom j sits in 32-byte records fetched with 256-bit loads.  No CPU fallback.
#include "eph_b200_atomic.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "eph_device.cuh"

using namespace ephb;

namespace epha {

constexpr int kLanes = 8;          // lanes per atom in the list sweeps
constexpr int kKappaShift = 16;    // bits 16..23 of the record's bit word: element index in the .kappa file

// EPH_Linear::reverse_lookup (eph_linear.h:50-60): std::upper_bound over the knots, then the inverse of the segment.
// Past the last knot the reference returns 0; below the first one it indexes knot -1 (undefined): 0 here.
__device__ __forceinline__ double lin_reverse(const double *__restrict__ y, int n, double dx, double yv) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (!(yv < y[mid])) lo = mid + 1; else hi = mid;
  }
  if (lo == n || lo == 0) return 0.0;
  const int idx = lo - 1;
  const double dy = (y[idx + 1] - y[idx]) / dx;
  return idx * dx + 1.0 / dy * (yv - y[idx]);
}

// EPH_Linear::operator() (eph_linear.h:40-47): truncating index; 0 beyond the table.  In the last interval the
// reference reads one element past its slope vector; the slope is taken as 0 there.
__device__ __forceinline__ double lin_eval(const double *__restrict__ y, int n, double dx, double x) {
  const double q = x / dx;
  if (!(q > -1.0) || !(q < static_cast<double>(n))) return 0.0;
  const int idx = static_cast<int>(q);
  const double dy = idx + 1 < n ? (y[idx + 1] - y[idx]) / dx : 0.0;
  return y[idx] + dy * (x - idx * dx);
}

struct Tables {
  const double2 *__restrict__ rho_tab;    // beta file: rho(r^2) [n_el][n_rho]{ab,cd}
  const double2 *__restrict__ alpha_tab;  // [n_el][n_beta]
  const double2 *__restrict__ beta_tab;
  int n_rho, n_beta;
  double inv_dr_sq, inv_drho, rc2, rho_cutoff;
  const double2 *__restrict__ rhoa_tab;   // kappa file: rho_a(r^2) [n_elk][n_r]
  int n_r, n_T;
  double inv_drk_sq, rk2, dT;
  const double *__restrict__ E_T;         // [n_elk][n_T]
  const double *__restrict__ K_T;         // [n_pairs][n_T], indexed by element
};

// {x, y, z, bits} {vx, vy, vz, 0} per atom; bits = beta element | group bit | kappa element << 16
__global__ void __launch_bounds__(256) pack_kernel(int nt, const double *__restrict__ x, const double *__restrict__ v,
                                                   const int *__restrict__ type, const int *__restrict__ mask,
                                                   const int *__restrict__ tmb, const int *__restrict__ tmk, int groupbit,
                                                   double4 *__restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  const int t = type[i] - 1;
  unsigned bits = (static_cast<unsigned>(tmb[t]) & kElemMask) | ((static_cast<unsigned>(tmk[t]) & 0xFFu) << kKappaShift);
  if (mask[i] & groupbit) bits |= kBitGroup;
  const size_t o = 3 * static_cast<size_t>(i);
  rec[2 * static_cast<size_t>(i)] = make_double4(x[o], x[o + 1], x[o + 2], bits_to_double(bits));
  rec[2 * static_cast<size_t>(i) + 1] = make_double4(v[o], v[o + 1], v[o + 2], 0.0);
}

// constructor, fix_eph_atomic.cpp:212-221
__global__ void __launch_bounds__(256) init_energy_kernel(int nlocal, const int *__restrict__ type, const int *__restrict__ mask,
                                                          const int *__restrict__ tmk, int groupbit, Tables t, double T_init,
                                                          double *__restrict__ E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double e = 0.0;
  if (mask[i] & groupbit) e = lin_eval(t.E_T + static_cast<size_t>(tmk[type[i] - 1]) * t.n_T, t.n_T, t.dT, T_init);
  E[i] = e;
}

// Comm::forward_comm(Fix*) on one rank: ghost g takes its owner's value (pack/unpack, fix_eph_atomic.cpp:849-927)
__global__ void __launch_bounds__(256) ghost_fill_kernel(int nlocal, int nghost, const int *__restrict__ owner,
                                                         double *__restrict__ a0, double *__restrict__ a1, double *__restrict__ a2) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int o = owner[g];
  if (a0) a0[nlocal + g] = a0[o];
  if (a1) a1[nlocal + g] = a1[o];
  if (a2) a2[nlocal + g] = a2[o];
}

// forward-comm payload of state RHO: {rho, rho_a} per atom (fix_eph_atomic.cpp:855-859, :893-897)
__global__ void __launch_bounds__(256) gather_kernel(int n, const int *__restrict__ list, int width, const double *__restrict__ a0,
                                                     const double *__restrict__ a1, double *__restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const size_t src = list[k];
  if (a1 != nullptr) { buf[2 * (size_t)k] = a0[src]; buf[2 * (size_t)k + 1] = a1[src]; }   // two scalars, interleaved
  else for (int d = 0; d < width; ++d) buf[(size_t)width * k + d] = a0[(size_t)width * src + d];
}

__global__ void __launch_bounds__(256) scatter_pair_kernel(int n, int first, const double *__restrict__ buf, double *__restrict__ a0,
                                                           double *__restrict__ a1) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  a0[first + k] = buf[2 * (size_t)k];
  a1[first + k] = buf[2 * (size_t)k + 1];
}

// xi_i for the group's local atoms (fix_eph_atomic.cpp:808-816): injected, or the counter-based stream keyed on the tag
__global__ void __launch_bounds__(256) xi_kernel(int nlocal, const int *__restrict__ mask, int groupbit, int do_random,
                                                 const double *__restrict__ inject, const long long *__restrict__ tag,
                                                 unsigned long long seed, unsigned long long step, double *__restrict__ xi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double r[3] = {0.0, 0.0, 0.0};
  if (do_random && (mask[i] & groupbit)) {
    if (inject) { r[0] = inject[3 * (size_t)i]; r[1] = inject[3 * (size_t)i + 1]; r[2] = inject[3 * (size_t)i + 2]; }
    else xi_stream(seed, step, tag[i], r);
  }
  xi[3 * (size_t)i] = r[0]; xi[3 * (size_t)i + 1] = r[1]; xi[3 * (size_t)i + 2] = r[2];
}

struct SweepArgs {
  int nlocal;
  const long long *__restrict__ offsets;
  const int *__restrict__ neigh;
  const double4 *__restrict__ rec;   // [nt][2]
  Tables t;
};

// calculate_environment (fix_eph_atomic.cpp:437-490): rho_i over r^2 < r_c^2 and the locality density rho_a_i over
// r^2 < r_kappa^2, neighbours outside the group skipped (:467); both 0 for atoms outside the group.
template <int LANES>
__global__ void __launch_bounds__(256) env_kernel(SweepArgs a, double *__restrict__ rho, double *__restrict__ rho_a) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int gpb = blockDim.x / LANES, gib = threadIdx.x / LANES;
  for (int i = blockIdx.x * gpb + gib; i < a.nlocal; i += gridDim.x * gpb) {
    const double4 pi = ld256(a.rec + 2 * (size_t)i);
    double r = 0.0, ra = 0.0;
    if (double_to_bits(pi.w) & kBitGroup) {
      const long long e = a.offsets[i + 1];
      for (long long k = a.offsets[i] + sub; k < e; k += LANES) {
        const int j = ld_stream(a.neigh + k) & kNeighMask;
        const double4 pj = ld256(a.rec + 2 * (size_t)j);
        const unsigned bj = double_to_bits(pj.w);
        if (!(bj & kBitGroup)) continue;
        const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
        const double r2 = ex * ex + ey * ey + ez * ez;
        if (r2 < a.t.rc2) r += spline_eval(a.t.rho_tab + 2 * (size_t)(bj & kElemMask) * a.t.n_rho, a.t.inv_dr_sq, r2);
        if (r2 < a.t.rk2) ra += spline_eval(a.t.rhoa_tab + 2 * (size_t)((bj >> kKappaShift) & 0xFFu) * a.t.n_r, a.t.inv_drk_sq, r2);
      }
      r = group_sum<LANES>(r, gmask);
      ra = group_sum<LANES>(ra, gmask);
    }
    if (sub == 0) { rho[i] = r; rho_a[i] = ra; }
  }
}

// per atom (locals and ghosts): cp = {s = alpha(rho)/rho, s sqrt(T(E)), rho > 0, T(E)}.  alpha once per atom instead
// of once per pair (fix_eph_atomic.cpp:517, :563, :575, :620, :637); T(E) = E_T_atomic.reverse(E) (:622, :639).
__global__ void __launch_bounds__(256) coupling_kernel(int nt, const double4 *__restrict__ rec, const double *__restrict__ rho,
                                                       const double *__restrict__ E, Tables t, double4 *__restrict__ cp) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nt) return;
  const unsigned bits = double_to_bits(rec[2 * (size_t)a].w);
  const double r = rho[a];
  double s = 0.0, sT = 0.0, T = 0.0, valid = 0.0;
  if (bits & kBitGroup) {
    T = lin_reverse(t.E_T + (size_t)((bits >> kKappaShift) & 0xFFu) * t.n_T, t.n_T, t.dT, E[a]);
    if (r > 0) {                        // `!(rho > 0)` guards, :515, :561, :618, :573, :635
      valid = 1.0;
      double alpha = 0.0;               // eph_beta.h:186-198
      if (!(r > t.rho_cutoff)) alpha = spline_eval(t.alpha_tab + 2 * (size_t)(bits & kElemMask) * t.n_beta, t.inv_drho, r);
      s = alpha / r;
      sT = sqrt(T) * s;
    }
  }
  cp[a] = make_double4(s, sT, valid, T);
}

// friction pass A (fix_eph_atomic.cpp:508-547): w_i = s_i sum_j [rho^{t_j}(r^2)/r^2] (e.(v_i - v_j)) e over group
// neighbours inside the cut-off; rho_j is not tested here.  0 for atoms outside the group or with rho_i <= 0.
template <int LANES>
__global__ void __launch_bounds__(256) w_kernel(SweepArgs a, const double4 *__restrict__ cp, int do_friction, double *__restrict__ w) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int gpb = blockDim.x / LANES, gib = threadIdx.x / LANES;
  for (int i = blockIdx.x * gpb + gib; i < a.nlocal; i += gridDim.x * gpb) {
    const double4 pi = ld256(a.rec + 2 * (size_t)i);
    const double4 ci = cp[i];
    double wx = 0.0, wy = 0.0, wz = 0.0;
    if (do_friction && (double_to_bits(pi.w) & kBitGroup) && ci.z != 0.0) {
      const double4 vi = ld256(a.rec + 2 * (size_t)i + 1);
      const long long e = a.offsets[i + 1];
      for (long long k = a.offsets[i] + sub; k < e; k += LANES) {
        const int j = ld_stream(a.neigh + k) & kNeighMask;
        const double4 pj = ld256(a.rec + 2 * (size_t)j);
        const unsigned bj = double_to_bits(pj.w);
        if (!(bj & kBitGroup)) continue;
        const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
        const double r2 = ex * ex + ey * ey + ez * ez;
        if (r2 >= a.t.rc2) continue;
        const double4 vj = ld256(a.rec + 2 * (size_t)j + 1);
        const double g = spline_eval(a.t.rho_tab + 2 * (size_t)(bj & kElemMask) * a.t.n_rho, a.t.inv_dr_sq, r2) / r2;
        const double d = g * (ex * (vi.x - vj.x) + ey * (vi.y - vj.y) + ez * (vi.z - vj.z));
        wx += d * ex; wy += d * ey; wz += d * ez;
      }
      wx = ci.x * group_sum<LANES>(wx, gmask);
      wy = ci.x * group_sum<LANES>(wy, gmask);
      wz = ci.x * group_sum<LANES>(wz, gmask);
    }
    if (sub == 0) { w[3 * (size_t)i] = wx; w[3 * (size_t)i + 1] = wy; w[3 * (size_t)i + 2] = wz; }
  }
}

// per atom (locals and ghosts; a ghost reads its owner's w and xi -- the WI and XI forward comms, :549-550, :818-819):
// q = {u = s w, valid} {z = s sqrt(T) xi, 0}
__global__ void __launch_bounds__(256) prep_kernel(int nlocal, int nt, const int *__restrict__ owner, const double4 *__restrict__ cp,
                                                   const double *__restrict__ w, const double *__restrict__ xi,
                                                   double4 *__restrict__ q) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nt) return;
  // owner == nullptr: the transport (LAMMPS' forward comm) has already filled the ghost rows of w and xi
  const int src = (a < nlocal || owner == nullptr) ? a : owner[a - nlocal];
  const double4 c = cp[a];
  const size_t o = 3 * (size_t)src;
  q[2 * (size_t)a] = make_double4(c.x * w[o], c.x * w[o + 1], c.x * w[o + 2], c.z);
  q[2 * (size_t)a + 1] = make_double4(c.y * xi[o], c.y * xi[o + 1], c.y * xi[o + 2], 0.0);
}

struct ForceArgs {
  const double4 *__restrict__ q;   // [nt][2]
  int do_friction, do_random, add_friction, add_random;   // FRICTION, RANDOM, and those without NOFRICTION / NORANDOM
  double eta_factor, dt;
  double *__restrict__ f;          // LAMMPS force array [nlocal][3] (read-modify-write) or nullptr
  double *__restrict__ f_eph, *__restrict__ f_rng, *__restrict__ dE;
};

// friction pass B and the random pass (fix_eph_atomic.cpp:554-606, :610-678) in one sweep: with g_ji = rho^{t_j}(r^2)/r^2,
// g_ij = rho^{t_i}(r^2)/r^2, over group neighbours inside the cut-off with rho_j > 0,
//   f_ij = [g_ji (e.u_i) - g_ij (e.u_j)] e          f_EPH_i -= f_ij      dE_i += 0.5 f_ij.(v_i - v_j) dt   (:599-603)
//   r_ij = eta [g_ji (e.z_i) - g_ij (e.z_j)] e      f_RNG_i += r_ij      dE_i -= 0.5 r_ij.(v_i - v_j) dt   (:671-675)
// and f += f_EPH (+ f_RNG) for group atoms (:830-848).
template <int LANES>
__global__ void __launch_bounds__(256) force_kernel(SweepArgs a, ForceArgs p) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int gpb = blockDim.x / LANES, gib = threadIdx.x / LANES;
  for (int i = blockIdx.x * gpb + gib; i < a.nlocal; i += gridDim.x * gpb) {
    const double4 pi = ld256(a.rec + 2 * (size_t)i);
    const unsigned bi = double_to_bits(pi.w);
    const double4 ui = ld256(p.q + 2 * (size_t)i);
    double fx = 0, fy = 0, fz = 0, rx = 0, ry = 0, rz = 0, de_f = 0, de_r = 0;
    const bool in_group = (bi & kBitGroup) != 0;
    if (in_group && ui.w != 0.0) {
      const double4 vi = ld256(a.rec + 2 * (size_t)i + 1);
      const double4 zi = ld256(p.q + 2 * (size_t)i + 1);
      const double2 *tab_i = a.t.rho_tab + 2 * (size_t)(bi & kElemMask) * a.t.n_rho;
      const long long e = a.offsets[i + 1];
      for (long long k = a.offsets[i] + sub; k < e; k += LANES) {
        const int j = ld_stream(a.neigh + k) & kNeighMask;
        const double4 pj = ld256(a.rec + 2 * (size_t)j);
        const unsigned bj = double_to_bits(pj.w);
        if (!(bj & kBitGroup)) continue;
        const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
        const double r2 = ex * ex + ey * ey + ez * ez;
        if (r2 >= a.t.rc2) continue;
        const double4 uj = ld256(p.q + 2 * (size_t)j);
        if (uj.w == 0.0) continue;                         // rho_j > 0 fails
        const double4 vj = ld256(a.rec + 2 * (size_t)j + 1);
        const double g_ji = spline_eval(a.t.rho_tab + 2 * (size_t)(bj & kElemMask) * a.t.n_rho, a.t.inv_dr_sq, r2) / r2;
        const double g_ij = ((bj ^ bi) & kElemMask) ? spline_eval(tab_i, a.t.inv_dr_sq, r2) / r2 : g_ji;
        const double ev = ex * (vi.x - vj.x) + ey * (vi.y - vj.y) + ez * (vi.z - vj.z);
        if (p.do_friction) {
          const double d = g_ji * (ex * ui.x + ey * ui.y + ez * ui.z) - g_ij * (ex * uj.x + ey * uj.y + ez * uj.z);
          fx -= d * ex; fy -= d * ey; fz -= d * ez;
          de_f += d * ev;
        }
        if (p.do_random) {
          const double4 zj = ld256(p.q + 2 * (size_t)j + 1);
          const double d = p.eta_factor * (g_ji * (ex * zi.x + ey * zi.y + ez * zi.z) - g_ij * (ex * zj.x + ey * zj.y + ez * zj.z));
          rx += d * ex; ry += d * ey; rz += d * ez;
          de_r += d * ev;
        }
      }
      fx = group_sum<LANES>(fx, gmask); fy = group_sum<LANES>(fy, gmask); fz = group_sum<LANES>(fz, gmask);
      rx = group_sum<LANES>(rx, gmask); ry = group_sum<LANES>(ry, gmask); rz = group_sum<LANES>(rz, gmask);
      de_f = group_sum<LANES>(de_f, gmask); de_r = group_sum<LANES>(de_r, gmask);
    }
    if (sub == 0) {
      const size_t o = 3 * (size_t)i;
      p.f_eph[o] = fx; p.f_eph[o + 1] = fy; p.f_eph[o + 2] = fz;
      p.f_rng[o] = rx; p.f_rng[o + 1] = ry; p.f_rng[o + 2] = rz;
      double de = 0.0;
      if (p.add_friction) de += 0.5 * de_f * p.dt;
      if (p.add_random) de -= 0.5 * de_r * p.dt;
      p.dE[i] = de;
      if (p.f != nullptr && in_group) {
        double ax = 0, ay = 0, az = 0;
        if (p.add_friction) { ax += fx; ay += fy; az += fz; }
        if (p.add_random) { ax += rx; ay += ry; az += rz; }
        p.f[o] += ax; p.f[o + 1] += ay; p.f[o + 2] += az;
      }
    }
  }
}

// heat_solve, first half of a loop (fix_eph_atomic.cpp:705-718): E_j += dE_j / loops, clamped at 0, for group atoms
__global__ void __launch_bounds__(256) heat_add_kernel(int nlocal, const double4 *__restrict__ rec, const double *__restrict__ dE,
                                                       double scaling, double *__restrict__ E) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nlocal) return;
  if (double_to_bits(rec[2 * (size_t)j].w) & kBitGroup) {
    double e = E[j] + dE[j] * scaling;
    if (e < 0.0) e = 0.0;
    E[j] = e;
  }
}

// the EI forward comm (:720-721) and the per-atom look-ups the reference repeats per pair (:729-731, :746-747):
// hk = {T(E), K(T), rho_a, 0} for locals and ghosts
__global__ void __launch_bounds__(256) heat_prep_kernel(int nlocal, int nt, const int *__restrict__ owner, const double4 *__restrict__ rec,
                                                        const double *__restrict__ rho_a, Tables t, double *__restrict__ E,
                                                        double4 *__restrict__ hk) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nt) return;
  double e;
  if (a < nlocal || owner == nullptr) e = E[a];   // owner == nullptr: ghost energies came through the transport
  else { e = E[owner[a - nlocal]]; E[a] = e; }
  const unsigned ek = (double_to_bits(rec[2 * (size_t)a].w) >> kKappaShift) & 0xFFu;
  const double T = lin_reverse(t.E_T + (size_t)ek * t.n_T, t.n_T, t.dT, e);
  const double K = lin_eval(t.K_T + (size_t)ek * t.n_T, t.n_T, t.dT, T);
  hk[a] = make_double4(T, K, rho_a[a], 0.0);
}

// heat_solve, second half of a loop (fix_eph_atomic.cpp:725-780): E1_j = E_j + 0.5 dt_loop sum_k K_jk (T_k - T_j)
// [rho_a^{t_k}(r^2)/rho_a_j + rho_a^{t_j}(r^2)/rho_a_k] over group neighbours with r^2 < r_kappa^2, clamped at 0
template <int LANES>
__global__ void __launch_bounds__(256) heat_kernel(SweepArgs a, const double4 *__restrict__ hk, const double *__restrict__ E,
                                                   double dt_loop, double *__restrict__ E1) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int gpb = blockDim.x / LANES, gib = threadIdx.x / LANES;
  for (int j = blockIdx.x * gpb + gib; j < a.nlocal; j += gridDim.x * gpb) {
    const double4 pj = ld256(a.rec + 2 * (size_t)j);
    const unsigned bj = double_to_bits(pj.w);
    double e1 = E[j];
    if (bj & kBitGroup) {
      const double4 hj = hk[j];
      const double2 *tab_j = a.t.rhoa_tab + 2 * (size_t)((bj >> kKappaShift) & 0xFFu) * a.t.n_r;
      double acc = 0.0;
      const long long e = a.offsets[j + 1];
      for (long long n = a.offsets[j] + sub; n < e; n += LANES) {
        const int k = ld_stream(a.neigh + n) & kNeighMask;
        const double4 pk = ld256(a.rec + 2 * (size_t)k);
        const unsigned bk = double_to_bits(pk.w);
        if (!(bk & kBitGroup)) continue;
        const double ex = pk.x - pj.x, ey = pk.y - pj.y, ez = pk.z - pj.z;
        const double r2 = ex * ex + ey * ey + ez * ez;
        if (r2 >= a.t.rk2) continue;
        const double4 hkk = ld256(hk + k);
        const double lK = 0.5 * (hj.y + hkk.y);
        const double dT = hkk.x - hj.x;
        const double v_rho_k = spline_eval(a.t.rhoa_tab + 2 * (size_t)((bk >> kKappaShift) & 0xFFu) * a.t.n_r, a.t.inv_drk_sq, r2);
        const double v_rho_j = ((bk ^ bj) >> kKappaShift) & 0xFFu ? spline_eval(tab_j, a.t.inv_drk_sq, r2) : v_rho_k;
        if (hj.z > 0.0) acc += lK * v_rho_k / hj.z * dT;
        if (hkk.z > 0.0) acc += lK * v_rho_j / hkk.z * dT;
      }
      acc = group_sum<LANES>(acc, gmask);
      e1 = e1 + 0.5 * acc * dt_loop;
      if (e1 < 0.0) e1 = 0.0;
    }
    if (sub == 0) E1[j] = e1;
  }
}

// end_of_step (fix_eph_atomic.cpp:370-380): T_a_i = T(E_i) for group atoms, sums of E, T and the atom count
__global__ void __launch_bounds__(256) summary_kernel(int nlocal, const int *__restrict__ type, const int *__restrict__ mask,
                                                      const int *__restrict__ tmk, int groupbit, Tables t,
                                                      const double *__restrict__ E, double *__restrict__ T_a, double *__restrict__ scal) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0, T = 0.0, c = 0.0;
  if (i < nlocal && (mask[i] & groupbit)) {
    e = E[i];
    T = lin_reverse(t.E_T + (size_t)tmk[type[i] - 1] * t.n_T, t.n_T, t.dT, e);
    T_a[i] = T;
    c = 1.0;
  }
  e = warp_sum(e); T = warp_sum(T); c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c != 0.0) {
    atomicAdd(scal + 0, e);
    atomicAdd(scal + 1, T);
    atomicAdd(scal + 2, c);
  }
}

// populate_array (fix_eph_atomic.cpp:401-435)
__global__ void __launch_bounds__(256) peratom_kernel(int nlocal, const int *__restrict__ type, const int *__restrict__ mask,
                                                      const int *__restrict__ tmb, int groupbit, Tables t, const double *__restrict__ rho,
                                                      const double *__restrict__ f_eph, const double *__restrict__ f_rng,
                                                      const double *__restrict__ rho_a, const double *__restrict__ E,
                                                      const double *__restrict__ dE, const double *__restrict__ T_a,
                                                      double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double r[12];
#pragma unroll
  for (int c = 0; c < 12; ++c) r[c] = 0.0;
  if (mask[i] & groupbit) {
    const double rh = rho[i];
    r[0] = rh;
    if (!(rh > t.rho_cutoff)) r[1] = spline_eval(t.beta_tab + 2 * (size_t)tmb[type[i] - 1] * t.n_beta, t.inv_drho, rh);  // eph_beta.h:171-184
    const size_t o = 3 * (size_t)i;
    r[2] = f_eph[o]; r[3] = f_eph[o + 1]; r[4] = f_eph[o + 2];
    r[5] = f_rng[o]; r[6] = f_rng[o + 1]; r[7] = f_rng[o + 2];
    r[8] = rho_a[i]; r[9] = E[i]; r[10] = dE[i]; r[11] = T_a[i];
  }
#pragma unroll
  for (int c = 0; c < 12; ++c) out[12 * (size_t)i + c] = r[c];
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  // grows to at least n elements; old content is kept and new storage is zeroed
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    const size_t want = n + n / 8 + 64;
    T *q = nullptr;
    // growth is rare (registration time): full device synchronisation on both sides keeps it ordered against the
    // engine's non-blocking stream, which the legacy-stream memset / copy below would otherwise not be
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&q, want * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemset(q, 0, want * sizeof(T));
    if (e == cudaSuccess && p) e = cudaMemcpy(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (p) cudaFree(p);
    p = q;
    cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

std::string g_create_error;

}  // namespace epha

using namespace epha;

struct eph_b200_atomic_handle {
  eph_b200_atomic_config cfg{};
  std::vector<int> tmb, tmk;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  std::string err;
  long long launches = 0;

  Tables t{};
  int n_elb = 0, n_elk = 0, n_pairs = 0;
  DevBuf<double2> rho_tab, alpha_tab, beta_tab, rhoa_tab;
  DevBuf<double> E_T, K_T;
  bool beta_set = false, kappa_set = false;
  double dt = 0, boltz = 0, eta = 0;
  bool dt_set = false;

  int nlocal = 0, nghost = 0;
  bool atoms_set = false, neigh_set = false, packed = false;
  DevBuf<int> type, mask, owner, d_tmb, d_tmk;
  DevBuf<long long> tag, off;
  DevBuf<int> neigh;
  const int *type_p = nullptr, *mask_p = nullptr;
  const long long *tag_p = nullptr, *off_p = nullptr;
  const int *neigh_p = nullptr;
  DevBuf<double> x, v, f, xi_in;
  // external ghost transport (LAMMPS' Comm::forward_comm(Fix*) with host buffers): no owner map, phase-split calls
  bool external_comm = false;
  int phase = 0;                 // 0 idle, 1 after post_force_begin, 2 after post_force_mid
  DevBuf<int> comm_idx;
  DevBuf<double> comm_buf;
  DevBuf<double4> rec, cp, q, hk;
  DevBuf<double> rho, rho_a, E, E1, dE, T_a, w, xi, f_eph, f_rng, array12, scal;
  double *h_pinned = nullptr;
};

namespace {

int fail(eph_b200_atomic_handle *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

#define EPHA_CUDA(h, call)                                                                                \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail(h, EPH_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define EPHA_LAUNCH_CHECK(h)                                                                              \
  do {                                                                                                    \
    ++(h)->launches;                                                                                      \
    cudaError_t e_ = cudaGetLastError();                                                                  \
    if (e_ != cudaSuccess)                                                                                \
      return fail(h, EPH_B200_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

inline int blocks_for(long long n, int threads) { return (int)std::max<long long>(1, (n + threads - 1) / threads); }

// grid of a list sweep: one resident wave (a multiple of the SM count), grid-stride over the atoms
inline int sweep_blocks(const eph_b200_atomic_handle *h) {
  const long long need = ((long long)h->nlocal * kLanes + 255) / 256;
  return (int)std::max<long long>(1, std::min<long long>(need, (long long)h->sm_count * 8));
}

template <class T>
int stage_in(eph_b200_atomic_handle *h, DevBuf<T> &buf, const T *src, size_t n, int memspace, const T **out) {
  if (memspace == EPH_B200_DEVICE) { *out = src; return EPH_B200_OK; }
  EPHA_CUDA(h, buf.reserve(std::max<size_t>(n, 1)));
  if (n) EPHA_CUDA(h, cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  *out = buf.p;
  return EPH_B200_OK;
}

SweepArgs sweep_args(const eph_b200_atomic_handle *h) {
  SweepArgs a;
  a.nlocal = h->nlocal;
  a.offsets = h->off_p;
  a.neigh = h->neigh_p;
  a.rec = h->rec.p;
  a.t = h->t;
  return a;
}

int ready(eph_b200_atomic_handle *h, const char *what) {
  if (!h->beta_set || !h->kappa_set) return fail(h, EPH_B200_ERR_ARG, "%s: tables not set", what);
  if (!h->dt_set) return fail(h, EPH_B200_ERR_ARG, "%s: set_dt not called", what);
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "%s: set_atoms not called", what);
  return EPH_B200_OK;
}

int run_summary(eph_b200_atomic_handle *h, double *Ee, double *Te) {
  const int nl = h->nlocal;
  EPHA_CUDA(h, cudaMemsetAsync(h->scal.p, 0, 4 * sizeof(double), h->stream));
  if (nl > 0) {
    summary_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(nl, h->type_p, h->mask_p, h->d_tmk.p, h->cfg.groupbit, h->t, h->E.p, h->T_a.p, h->scal.p);
    EPHA_LAUNCH_CHECK(h);
  }
  if (Ee || Te) {
    EPHA_CUDA(h, cudaMemcpyAsync(h->h_pinned, h->scal.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
    if (Ee) *Ee = h->h_pinned[0];
    // T_local / atom_counter, then / proc_counter (1 on one rank; 0 atoms in the group divide by zero as in :387-397)
    if (Te) *Te = h->h_pinned[2] > 0.0 ? h->h_pinned[1] / h->h_pinned[2] : std::nan("");
  }
  return EPH_B200_OK;
}

}  // namespace

extern "C" {

const char *eph_b200_atomic_create_error(void) { return g_create_error.c_str(); }
const char *eph_b200_atomic_last_error(const eph_b200_atomic_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }
long long eph_b200_atomic_launch_count(const eph_b200_atomic_handle *h) { return h ? h->launches : 0; }

int eph_b200_atomic_create(const eph_b200_atomic_config *cfg, eph_b200_atomic_handle **out) {
  if (!cfg || !out) { g_create_error = "eph_b200_atomic_create: null argument"; return EPH_B200_ERR_ARG; }
  *out = nullptr;
  if (cfg->ntypes < 1 || !cfg->type_map_beta || !cfg->type_map_kappa) {
    g_create_error = "eph_b200_atomic_create: ntypes < 1 or a type map is missing";
    return EPH_B200_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1) {
    g_create_error = std::string("eph_b200_atomic_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    return EPH_B200_ERR_NODEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "eph_b200_atomic_create: bad device ordinal"; return EPH_B200_ERR_ARG; }
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) {
    g_create_error = std::string("eph_b200_atomic_create: ") + cudaGetErrorString(e);
    return EPH_B200_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "eph_b200_atomic_create: kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor);
    return EPH_B200_ERR_NODEVICE;
  }
  auto *h = new eph_b200_atomic_handle;
  h->cfg = *cfg;
  h->tmb.assign(cfg->type_map_beta, cfg->type_map_beta + cfg->ntypes);
  h->tmk.assign(cfg->type_map_kappa, cfg->type_map_kappa + cfg->ntypes);
  h->cfg.type_map_beta = h->tmb.data();
  h->cfg.type_map_kappa = h->tmk.data();
  h->sm_count = prop.multiProcessorCount;
  if (cfg->stream) h->stream = static_cast<cudaStream_t>(cfg->stream);
  else {
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
      g_create_error = std::string("eph_b200_atomic_create: ") + cudaGetErrorString(e);
      delete h;
      return EPH_B200_ERR_CUDA;
    }
    h->own_stream = true;
  }
  const bool ok = h->d_tmb.reserve(cfg->ntypes) == cudaSuccess && h->d_tmk.reserve(cfg->ntypes) == cudaSuccess &&
                  h->scal.reserve(4) == cudaSuccess && cudaMallocHost(&h->h_pinned, 4 * sizeof(double)) == cudaSuccess &&
                  cudaMemcpy(h->d_tmb.p, h->tmb.data(), cfg->ntypes * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(h->d_tmk.p, h->tmk.data(), cfg->ntypes * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) {
    g_create_error = std::string("eph_b200_atomic_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    eph_b200_atomic_destroy(h);
    return EPH_B200_ERR_CUDA;
  }
  *out = h;
  return EPH_B200_OK;
}

int eph_b200_atomic_destroy(eph_b200_atomic_handle *h) {
  if (!h) return EPH_B200_OK;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  h->rho_tab.release(); h->alpha_tab.release(); h->beta_tab.release(); h->rhoa_tab.release(); h->E_T.release(); h->K_T.release();
  h->type.release(); h->mask.release(); h->owner.release(); h->d_tmb.release(); h->d_tmk.release(); h->tag.release();
  h->off.release(); h->neigh.release(); h->x.release(); h->v.release(); h->f.release(); h->xi_in.release(); h->comm_idx.release(); h->comm_buf.release();
  h->rec.release(); h->cp.release(); h->q.release(); h->hk.release();
  h->rho.release(); h->rho_a.release(); h->E.release(); h->E1.release(); h->dE.release(); h->T_a.release(); h->w.release();
  h->xi.release(); h->f_eph.release(); h->f_rng.release(); h->array12.release(); h->scal.release();
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return EPH_B200_OK;
}

int eph_b200_atomic_synchronize(eph_b200_atomic_handle *h) {
  if (!h) return EPH_B200_ERR_ARG;
  EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPH_B200_OK;
}

int eph_b200_atomic_set_beta_tables(eph_b200_atomic_handle *h, int n_elements, int n_rho, double inv_dr_sq,
                                    const double *coeff_rho_r_sq, int n_beta, double inv_drho, const double *coeff_alpha,
                                    const double *coeff_beta, double r_cutoff_sq, double rho_cutoff) {
  if (!h) return EPH_B200_ERR_ARG;
  if (n_elements < 1 || n_elements > 255 || n_rho < 4 || n_beta < 4 || !coeff_rho_r_sq || !coeff_alpha || !coeff_beta)
    return fail(h, EPH_B200_ERR_ARG, "atomic_set_beta_tables: bad table sizes or null coefficients");
  for (int t = 0; t < h->cfg.ntypes; ++t)
    if (h->tmb[t] < 0 || h->tmb[t] >= n_elements)
      return fail(h, EPH_B200_ERR_ARG, "atomic_set_beta_tables: type %d maps to element %d, table has %d", t + 1, h->tmb[t], n_elements);
  cudaSetDevice(h->cfg.device);
  const size_t nr = (size_t)n_elements * n_rho * 2, nb = (size_t)n_elements * n_beta * 2;
  EPHA_CUDA(h, h->rho_tab.reserve(nr));
  EPHA_CUDA(h, h->alpha_tab.reserve(nb));
  EPHA_CUDA(h, h->beta_tab.reserve(nb));
  EPHA_CUDA(h, cudaMemcpyAsync(h->rho_tab.p, coeff_rho_r_sq, nr * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  EPHA_CUDA(h, cudaMemcpyAsync(h->alpha_tab.p, coeff_alpha, nb * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  EPHA_CUDA(h, cudaMemcpyAsync(h->beta_tab.p, coeff_beta, nb * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_elb = n_elements;
  h->t.rho_tab = h->rho_tab.p; h->t.alpha_tab = h->alpha_tab.p; h->t.beta_tab = h->beta_tab.p;
  h->t.n_rho = n_rho; h->t.n_beta = n_beta; h->t.inv_dr_sq = inv_dr_sq; h->t.inv_drho = inv_drho;
  h->t.rc2 = r_cutoff_sq; h->t.rho_cutoff = rho_cutoff;
  h->beta_set = true;
  return EPH_B200_OK;
}

int eph_b200_atomic_set_kappa_tables(eph_b200_atomic_handle *h, int n_elements, int n_pairs, int n_r, double inv_dr_sq,
                                     const double *coeff_rho_r_sq, double r_cutoff_sq, int n_T, double dT,
                                     const double *E_T, const double *K_T) {
  if (!h) return EPH_B200_ERR_ARG;
  if (n_elements < 1 || n_elements > 255 || n_r < 4 || n_T < 2 || !(dT > 0) || !coeff_rho_r_sq || !E_T || !K_T)
    return fail(h, EPH_B200_ERR_ARG, "atomic_set_kappa_tables: bad table sizes or null tables");
  if (n_pairs < n_elements)
    return fail(h, EPH_B200_ERR_ARG, "atomic_set_kappa_tables: K(T) is indexed by element (fix_eph_atomic.cpp:731) but the file "
                                     "holds %d table(s) for %d elements (eph_kappa.h:69)", n_pairs, n_elements);
  for (int t = 0; t < h->cfg.ntypes; ++t)
    if (h->tmk[t] < 0 || h->tmk[t] >= n_elements)
      return fail(h, EPH_B200_ERR_ARG, "atomic_set_kappa_tables: type %d maps to element %d, table has %d", t + 1, h->tmk[t], n_elements);
  cudaSetDevice(h->cfg.device);
  const size_t nr = (size_t)n_elements * n_r * 2;
  EPHA_CUDA(h, h->rhoa_tab.reserve(nr));
  EPHA_CUDA(h, h->E_T.reserve((size_t)n_elements * n_T));
  EPHA_CUDA(h, h->K_T.reserve((size_t)n_pairs * n_T));
  EPHA_CUDA(h, cudaMemcpyAsync(h->rhoa_tab.p, coeff_rho_r_sq, nr * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  EPHA_CUDA(h, cudaMemcpyAsync(h->E_T.p, E_T, (size_t)n_elements * n_T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPHA_CUDA(h, cudaMemcpyAsync(h->K_T.p, K_T, (size_t)n_pairs * n_T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_elk = n_elements; h->n_pairs = n_pairs;
  h->t.rhoa_tab = h->rhoa_tab.p; h->t.n_r = n_r; h->t.inv_drk_sq = inv_dr_sq; h->t.rk2 = r_cutoff_sq;
  h->t.n_T = n_T; h->t.dT = dT; h->t.E_T = h->E_T.p; h->t.K_T = h->K_T.p;
  h->kappa_set = true;
  return EPH_B200_OK;
}

int eph_b200_atomic_set_dt(eph_b200_atomic_handle *h, double dt, double boltz) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!(dt > 0) || !(boltz > 0)) return fail(h, EPH_B200_ERR_ARG, "atomic_set_dt: dt and boltz must be positive");
  h->dt = dt; h->boltz = boltz;
  h->eta = std::sqrt(2.0 * boltz / dt);   // fix_eph_atomic.cpp:810
  h->dt_set = true;
  return EPH_B200_OK;
}

int eph_b200_atomic_set_atoms(eph_b200_atomic_handle *h, int nlocal, int nghost, const int *type, const int *mask,
                              const int64_t *tag, const int *ghost_owner, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (nlocal < 0 || nghost < 0 || !type || !mask || !tag)
    return fail(h, EPH_B200_ERR_ARG, "atomic_set_atoms: bad counts or null arrays");
  if (nghost > 0 && !ghost_owner && !h->external_comm)
    return fail(h, EPH_B200_ERR_ARG, "atomic_set_atoms: every ghost needs an owner on this rank (or set_comm_mode(external))");
  cudaSetDevice(h->cfg.device);
  const size_t nt = (size_t)nlocal + nghost;
  if (memspace == EPH_B200_HOST) {
    for (size_t i = 0; i < nt; ++i)
      if (type[i] < 1 || type[i] > h->cfg.ntypes) return fail(h, EPH_B200_ERR_ARG, "atomic_set_atoms: atom %zu has type %d", i, type[i]);
    for (int g = 0; g < nghost && ghost_owner && !h->external_comm; ++g)
      if (ghost_owner[g] < 0 || ghost_owner[g] >= nlocal)
        return fail(h, EPH_B200_ERR_ARG, "atomic_set_atoms: ghost %d has no owner on this rank", g);
  }
  int rc;
  const long long *tag_ll = reinterpret_cast<const long long *>(tag);
  if ((rc = stage_in(h, h->type, type, nt, memspace, &h->type_p)) != EPH_B200_OK) return rc;
  if ((rc = stage_in(h, h->mask, mask, nt, memspace, &h->mask_p)) != EPH_B200_OK) return rc;
  if ((rc = stage_in(h, h->tag, tag_ll, nt, memspace, &h->tag_p)) != EPH_B200_OK) return rc;
  // the owner map is always copied: it is read by later calls
  EPHA_CUDA(h, h->owner.reserve(std::max<size_t>(nghost, 1)));
  if (nghost && ghost_owner) EPHA_CUDA(h, cudaMemcpyAsync(h->owner.p, ghost_owner, (size_t)nghost * sizeof(int),
                                           memspace == EPH_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  const size_t n1 = std::max<size_t>(nt, 1), nl1 = std::max<size_t>(nlocal, 1);
  EPHA_CUDA(h, h->rec.reserve(2 * n1)); EPHA_CUDA(h, h->cp.reserve(n1)); EPHA_CUDA(h, h->q.reserve(2 * n1)); EPHA_CUDA(h, h->hk.reserve(n1));
  EPHA_CUDA(h, h->rho.reserve(n1)); EPHA_CUDA(h, h->rho_a.reserve(n1)); EPHA_CUDA(h, h->E.reserve(n1));
  EPHA_CUDA(h, h->E1.reserve(nl1)); EPHA_CUDA(h, h->dE.reserve(nl1)); EPHA_CUDA(h, h->T_a.reserve(nl1));
  EPHA_CUDA(h, h->w.reserve(3 * n1)); EPHA_CUDA(h, h->xi.reserve(3 * n1)); EPHA_CUDA(h, h->f_eph.reserve(3 * nl1));
  EPHA_CUDA(h, h->f_rng.reserve(3 * nl1)); EPHA_CUDA(h, h->array12.reserve(12 * nl1));
  EPHA_CUDA(h, cudaStreamSynchronize(h->stream));   // host arrays may be reused by the caller
  h->nlocal = nlocal; h->nghost = nghost;
  h->atoms_set = true;
  h->neigh_set = false;
  h->packed = false;
  h->phase = 0;
  return EPH_B200_OK;
}

int eph_b200_atomic_set_neighbors_csr(eph_b200_atomic_handle *h, int nlocal, const int64_t *offsets, const int *neigh, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->atoms_set || nlocal != h->nlocal || !offsets) return fail(h, EPH_B200_ERR_ARG, "atomic_set_neighbors: set_atoms first, same nlocal");
  cudaSetDevice(h->cfg.device);
  const long long *off_ll = reinterpret_cast<const long long *>(offsets);
  long long n_entries = 0;
  if (memspace == EPH_B200_HOST) {
    n_entries = offsets[nlocal];
    const long long nt = (long long)h->nlocal + h->nghost;
    for (int i = 0; i < nlocal; ++i)
      if (offsets[i + 1] < offsets[i]) return fail(h, EPH_B200_ERR_ARG, "atomic_set_neighbors: offsets not monotonic at %d", i);
    for (long long k = 0; k < n_entries; ++k)
      if ((neigh[k] & kNeighMask) >= nt) return fail(h, EPH_B200_ERR_ARG, "atomic_set_neighbors: entry %lld out of range", k);
  } else {
    EPHA_CUDA(h, cudaMemcpyAsync(&n_entries, off_ll + nlocal, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  if (n_entries > 0 && !neigh) return fail(h, EPH_B200_ERR_ARG, "atomic_set_neighbors: null list");
  int rc;
  if ((rc = stage_in(h, h->off, off_ll, (size_t)nlocal + 1, memspace, &h->off_p)) != EPH_B200_OK) return rc;
  if ((rc = stage_in(h, h->neigh, neigh, (size_t)n_entries, memspace, &h->neigh_p)) != EPH_B200_OK) return rc;
  EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  h->neigh_set = true;
  return EPH_B200_OK;
}

int eph_b200_atomic_init_energy(eph_b200_atomic_handle *h, double T_init) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->kappa_set || !h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "atomic_init_energy: kappa tables and atoms first");
  cudaSetDevice(h->cfg.device);
  if (h->nlocal > 0) {
    init_energy_kernel<<<blocks_for(h->nlocal, 256), 256, 0, h->stream>>>(h->nlocal, h->type_p, h->mask_p, h->d_tmk.p, h->cfg.groupbit, h->t, T_init, h->E.p);
    EPHA_LAUNCH_CHECK(h);
  }
  return EPH_B200_OK;
}

int eph_b200_atomic_set_energy(eph_b200_atomic_handle *h, const double *E, int memspace) {
  if (!h || !E) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "atomic_set_energy: set_atoms first");
  cudaSetDevice(h->cfg.device);
  if (h->nlocal > 0) {
    EPHA_CUDA(h, cudaMemcpyAsync(h->E.p, E, (size_t)h->nlocal * sizeof(double),
                                 memspace == EPH_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    if (memspace == EPH_B200_HOST) EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

int eph_b200_atomic_get_energy(eph_b200_atomic_handle *h, double *E, int memspace) {
  if (!h || !E) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "atomic_get_energy: set_atoms first");
  cudaSetDevice(h->cfg.device);
  if (h->nlocal > 0) {
    EPHA_CUDA(h, cudaMemcpyAsync(E, h->E.p, (size_t)h->nlocal * sizeof(double),
                                 memspace == EPH_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    if (memspace == EPH_B200_HOST) EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

int eph_b200_atomic_set_comm_mode(eph_b200_atomic_handle *h, int external) {
  if (!h) return EPH_B200_ERR_ARG;
  h->external_comm = external != 0;
  h->phase = 0;
  return EPH_B200_OK;
}

// post_force in three parts; between them the ghost rows are filled either internally through the owner map
// (eph_b200_atomic_post_force) or by the caller's transport through pack/unpack_forward (LAMMPS' forward comm).
int eph_b200_atomic_post_force_begin(eph_b200_atomic_handle *h, const double *x, const double *v, const double *xi_inject,
                                     long long ntimestep, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  int rc = ready(h, "atomic_post_force");
  if (rc != EPH_B200_OK) return rc;
  if (!h->neigh_set) return fail(h, EPH_B200_ERR_ARG, "atomic_post_force: set_neighbors not called");
  if (!x || !v) return fail(h, EPH_B200_ERR_ARG, "atomic_post_force: null x or v");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal, ng = h->nghost, nt = nl + ng;
  h->phase = 1;
  if (nt == 0) return EPH_B200_OK;
  const bool do_random = (h->cfg.flags & EPH_B200_RANDOM) != 0;
  const double *xd, *vd, *xid = nullptr;
  if ((rc = stage_in(h, h->x, x, 3 * (size_t)nt, memspace, &xd)) != EPH_B200_OK) return rc;
  if ((rc = stage_in(h, h->v, v, 3 * (size_t)nt, memspace, &vd)) != EPH_B200_OK) return rc;
  if (xi_inject && do_random && (rc = stage_in(h, h->xi_in, xi_inject, 3 * (size_t)nl, memspace, &xid)) != EPH_B200_OK) return rc;
  cudaStream_t st = h->stream;
  pack_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(nt, xd, vd, h->type_p, h->mask_p, h->d_tmb.p, h->d_tmk.p, h->cfg.groupbit, h->rec.p);
  EPHA_LAUNCH_CHECK(h);
  h->packed = true;
  if (nl > 0) {
    xi_kernel<<<blocks_for(nl, 256), 256, 0, st>>>(nl, h->mask_p, h->cfg.groupbit, do_random ? 1 : 0, xid, h->tag_p, h->cfg.seed, (unsigned long long)ntimestep, h->xi.p);
    EPHA_LAUNCH_CHECK(h);
    const SweepArgs sa = sweep_args(h);
    env_kernel<kLanes><<<sweep_blocks(h), 256, 0, st>>>(sa, h->rho.p, h->rho_a.p);
    EPHA_LAUNCH_CHECK(h);
  }
  return EPH_B200_OK;
}

// needs rho, rho_a and E of the ghosts (forward comms RHO and EI, fix_eph_atomic.cpp:803-804, :824-825)
int eph_b200_atomic_post_force_mid(eph_b200_atomic_handle *h) {
  if (!h) return EPH_B200_ERR_ARG;
  if (h->phase != 1) return fail(h, EPH_B200_ERR_ARG, "atomic_post_force_mid without post_force_begin");
  cudaSetDevice(h->cfg.device);
  h->phase = 2;
  const int nl = h->nlocal, nt = nl + h->nghost;
  if (nt == 0) return EPH_B200_OK;
  cudaStream_t st = h->stream;
  coupling_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(nt, h->rec.p, h->rho.p, h->E.p, h->t, h->cp.p);
  EPHA_LAUNCH_CHECK(h);
  if (nl > 0) {
    const SweepArgs sa = sweep_args(h);
    w_kernel<kLanes><<<sweep_blocks(h), 256, 0, st>>>(sa, h->cp.p, (h->cfg.flags & EPH_B200_FRICTION) ? 1 : 0, h->w.p);
    EPHA_LAUNCH_CHECK(h);
  }
  return EPH_B200_OK;
}

// needs w and xi of the ghosts (forward comms WI and XI, :549-550, :818-819): through the owner map, or already in place
int eph_b200_atomic_post_force_end(eph_b200_atomic_handle *h, double *f, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (h->phase != 2) return fail(h, EPH_B200_ERR_ARG, "atomic_post_force_end without post_force_mid");
  cudaSetDevice(h->cfg.device);
  h->phase = 0;
  const int nl = h->nlocal, nt = nl + h->nghost;
  if (nt == 0) return EPH_B200_OK;
  const int flags = h->cfg.flags;
  const bool do_friction = (flags & EPH_B200_FRICTION) != 0, do_random = (flags & EPH_B200_RANDOM) != 0;
  cudaStream_t st = h->stream;
  double *fd = f;
  if (f && memspace == EPH_B200_HOST) {
    EPHA_CUDA(h, h->f.reserve(3 * (size_t)std::max(nl, 1)));
    EPHA_CUDA(h, cudaMemcpyAsync(h->f.p, f, 3 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, st));
    fd = h->f.p;
  }
  prep_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(nl, nt, h->external_comm ? nullptr : h->owner.p, h->cp.p, h->w.p, h->xi.p, h->q.p);
  EPHA_LAUNCH_CHECK(h);
  if (nl > 0) {
    ForceArgs p;
    p.q = h->q.p;
    p.do_friction = do_friction ? 1 : 0;
    p.do_random = do_random ? 1 : 0;
    p.add_friction = (do_friction && !(flags & EPH_B200_NOFRICTION)) ? 1 : 0;
    p.add_random = (do_random && !(flags & EPH_B200_NORANDOM)) ? 1 : 0;
    p.eta_factor = h->eta;
    p.dt = h->dt;
    p.f = fd;
    p.f_eph = h->f_eph.p; p.f_rng = h->f_rng.p; p.dE = h->dE.p;
    const SweepArgs sa = sweep_args(h);
    force_kernel<kLanes><<<sweep_blocks(h), 256, 0, st>>>(sa, p);
    EPHA_LAUNCH_CHECK(h);
  }
  if (f && memspace == EPH_B200_HOST) {
    EPHA_CUDA(h, cudaMemcpyAsync(f, h->f.p, 3 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, st));
    EPHA_CUDA(h, cudaStreamSynchronize(st));
  }
  return EPH_B200_OK;
}

int eph_b200_atomic_post_force(eph_b200_atomic_handle *h, const double *x, const double *v, double *f, const double *xi_inject,
                               long long ntimestep, int memspace) {
  if (!h) return EPH_B200_ERR_ARG;
  if (h->external_comm) return fail(h, EPH_B200_ERR_ARG, "atomic_post_force: external comm mode uses post_force_begin / mid / end");
  int rc = eph_b200_atomic_post_force_begin(h, x, v, xi_inject, ntimestep, memspace);
  if (rc != EPH_B200_OK) return rc;
  if (h->nghost > 0) {   // the EI and RHO forward comms (:803-804, :824-825) through the owner map
    ghost_fill_kernel<<<blocks_for(h->nghost, 256), 256, 0, h->stream>>>(h->nlocal, h->nghost, h->owner.p, h->rho.p, h->rho_a.p, h->E.p);
    EPHA_LAUNCH_CHECK(h);
  }
  if ((rc = eph_b200_atomic_post_force_mid(h)) != EPH_B200_OK) return rc;
  return eph_b200_atomic_post_force_end(h, f, memspace);   // prep reads the ghosts' w and xi from their owners
}

// Comm::forward_comm(Fix*) with host buffers: FixEPHAtomic::pack_forward_comm / unpack_forward_comm
// (fix_eph_atomic.cpp:849-927).  state: 1 RHO {rho, rho_a}, 2 XI, 3 WI (3 doubles), 4 EI (1 double).
int eph_b200_atomic_pack_forward(eph_b200_atomic_handle *h, int state, int n, const int *list, double *buf) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->atoms_set || n < 0 || (n > 0 && (!list || !buf))) return fail(h, EPH_B200_ERR_ARG, "atomic_pack_forward: bad arguments");
  const int width = state == 1 ? 2 : state == 4 ? 1 : (state == 2 || state == 3) ? 3 : 0;
  if (width == 0) return fail(h, EPH_B200_ERR_ARG, "atomic_pack_forward: unknown state %d", state);
  if (n == 0) return 0;
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = h->stream;
  EPHA_CUDA(h, h->comm_idx.reserve(n));
  EPHA_CUDA(h, h->comm_buf.reserve((size_t)width * n));
  EPHA_CUDA(h, cudaMemcpyAsync(h->comm_idx.p, list, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  const double *a0 = state == 1 ? h->rho.p : state == 2 ? h->xi.p : state == 3 ? h->w.p : h->E.p;
  const double *a1 = state == 1 ? h->rho_a.p : nullptr;
  gather_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, h->comm_idx.p, width, a0, a1, h->comm_buf.p);
  EPHA_LAUNCH_CHECK(h);
  EPHA_CUDA(h, cudaMemcpyAsync(buf, h->comm_buf.p, (size_t)width * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  EPHA_CUDA(h, cudaStreamSynchronize(st));
  return width * n;
}

int eph_b200_atomic_unpack_forward(eph_b200_atomic_handle *h, int state, int n, int first, const double *buf) {
  if (!h) return EPH_B200_ERR_ARG;
  const long long nt = (long long)h->nlocal + h->nghost;
  if (!h->atoms_set || n < 0 || first < 0 || (long long)first + n > nt || (n > 0 && !buf))
    return fail(h, EPH_B200_ERR_ARG, "atomic_unpack_forward: bad arguments");
  if (n == 0) return EPH_B200_OK;
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = h->stream;
  if (state == 1) {
    EPHA_CUDA(h, h->comm_buf.reserve(2 * (size_t)n));
    EPHA_CUDA(h, cudaMemcpyAsync(h->comm_buf.p, buf, 2 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    scatter_pair_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, first, h->comm_buf.p, h->rho.p, h->rho_a.p);
    EPHA_LAUNCH_CHECK(h);
  } else if (state == 2 || state == 3) {   // rows [first, first + n) are contiguous
    double *dst = (state == 2 ? h->xi.p : h->w.p) + 3 * (size_t)first;
    EPHA_CUDA(h, cudaMemcpyAsync(dst, buf, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  } else if (state == 4) {
    EPHA_CUDA(h, cudaMemcpyAsync(h->E.p + first, buf, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  } else {
    return fail(h, EPH_B200_ERR_ARG, "atomic_unpack_forward: unknown state %d", state);
  }
  EPHA_CUDA(h, cudaStreamSynchronize(st));   // the caller's buffer is reused for the next swap
  return EPH_B200_OK;
}

int eph_b200_atomic_summary(eph_b200_atomic_handle *h, double *Ee, double *Te) {
  if (!h) return EPH_B200_ERR_ARG;
  if (!h->kappa_set || !h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "atomic_summary: kappa tables and atoms first");
  cudaSetDevice(h->cfg.device);
  return run_summary(h, Ee, Te);
}

// heat_solve (fix_eph_atomic.cpp:681-787) loop by loop: heat_begin adds the loop's share of the ledger (:705-718), then
// the ghosts' energies are refreshed (forward comm EI, :720-721: internally, or by the caller), heat_end diffuses (:725-785)
int eph_b200_atomic_heat_loops(const eph_b200_atomic_handle *h) {
  if (!h || !(h->cfg.flags & EPH_B200_FDM)) return 0;
  return h->cfg.inner_loops > 0 ? h->cfg.inner_loops : 1;   // :695-699
}

int eph_b200_atomic_heat_begin(eph_b200_atomic_handle *h) {
  if (!h) return EPH_B200_ERR_ARG;
  int rc = ready(h, "atomic_heat_begin");
  if (rc != EPH_B200_OK) return rc;
  if (!h->packed || !h->neigh_set) return fail(h, EPH_B200_ERR_ARG, "atomic_heat_begin: no post_force since the atoms were registered");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  if (nl == 0) return EPH_B200_OK;
  const double scaling = 1.0 / static_cast<double>(eph_b200_atomic_heat_loops(h) > 0 ? eph_b200_atomic_heat_loops(h) : 1);
  heat_add_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(nl, h->rec.p, h->dE.p, scaling, h->E.p);
  EPHA_LAUNCH_CHECK(h);
  return EPH_B200_OK;
}

int eph_b200_atomic_heat_end(eph_b200_atomic_handle *h) {
  if (!h) return EPH_B200_ERR_ARG;
  int rc = ready(h, "atomic_heat_end");
  if (rc != EPH_B200_OK) return rc;
  if (!h->packed || !h->neigh_set) return fail(h, EPH_B200_ERR_ARG, "atomic_heat_end: no post_force since the atoms were registered");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal, nt = nl + h->nghost;
  if (nl == 0) return EPH_B200_OK;
  cudaStream_t st = h->stream;
  const int loops = eph_b200_atomic_heat_loops(h) > 0 ? eph_b200_atomic_heat_loops(h) : 1;
  const double dt_loop = h->dt * (1.0 / static_cast<double>(loops));
  const SweepArgs sa = sweep_args(h);
  heat_prep_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(nl, nt, h->external_comm ? nullptr : h->owner.p, h->rec.p, h->rho_a.p, h->t, h->E.p, h->hk.p);
  EPHA_LAUNCH_CHECK(h);
  heat_kernel<kLanes><<<sweep_blocks(h), 256, 0, st>>>(sa, h->hk.p, h->E.p, dt_loop, h->E1.p);
  EPHA_LAUNCH_CHECK(h);
  EPHA_CUDA(h, cudaMemcpyAsync(h->E.p, h->E1.p, (size_t)nl * sizeof(double), cudaMemcpyDeviceToDevice, st));   // :783-785
  return EPH_B200_OK;
}

int eph_b200_atomic_end_of_step(eph_b200_atomic_handle *h, double *Ee, double *Te) {
  if (!h) return EPH_B200_ERR_ARG;
  int rc = ready(h, "atomic_end_of_step");
  if (rc != EPH_B200_OK) return rc;
  if (h->external_comm && eph_b200_atomic_heat_loops(h) > 0)
    return fail(h, EPH_B200_ERR_ARG, "atomic_end_of_step: external comm mode uses heat_begin / heat_end per loop, then summary");
  cudaSetDevice(h->cfg.device);
  if (h->nlocal > 0)
    for (int it = 0; it < eph_b200_atomic_heat_loops(h); ++it) {   // Flag::HEAT -> heat_solve, fix_eph_atomic.cpp:370
      if ((rc = eph_b200_atomic_heat_begin(h)) != EPH_B200_OK) return rc;
      if ((rc = eph_b200_atomic_heat_end(h)) != EPH_B200_OK) return rc;     // heat_prep copies the owners' energies to the ghosts
    }
  return run_summary(h, Ee, Te);
}

int eph_b200_atomic_get_peratom(eph_b200_atomic_handle *h, double *array12, int memspace) {
  if (!h || !array12) return EPH_B200_ERR_ARG;
  if (!h->beta_set || !h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "atomic_get_peratom: tables and atoms first");
  cudaSetDevice(h->cfg.device);
  const int nl = h->nlocal;
  if (nl == 0) return EPH_B200_OK;
  double *dst = memspace == EPH_B200_DEVICE ? array12 : h->array12.p;
  peratom_kernel<<<blocks_for(nl, 256), 256, 0, h->stream>>>(nl, h->type_p, h->mask_p, h->d_tmb.p, h->cfg.groupbit, h->t, h->rho.p, h->f_eph.p, h->f_rng.p, h->rho_a.p, h->E.p, h->dE.p, h->T_a.p, dst);
  EPHA_LAUNCH_CHECK(h);
  if (memspace == EPH_B200_HOST) {
    EPHA_CUDA(h, cudaMemcpyAsync(array12, dst, 12 * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

int eph_b200_atomic_get_probe(eph_b200_atomic_handle *h, int which, double *out) {
  if (!h || !out) return EPH_B200_ERR_ARG;
  if (!h->atoms_set) return fail(h, EPH_B200_ERR_ARG, "atomic_get_probe: set_atoms first");
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
    case 5: src = h->rho_a.p; n = nt; break;
    case 6: src = h->E.p; n = nt; break;
    case 7: src = h->dE.p; n = nl; break;
    case 8: src = h->T_a.p; n = nl; break;
    default: return fail(h, EPH_B200_ERR_ARG, "atomic_get_probe: unknown probe %d", which);
  }
  if (n) {
    EPHA_CUDA(h, cudaMemcpyAsync(out, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPHA_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return EPH_B200_OK;
}

}  // extern "C"
