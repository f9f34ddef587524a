// Device-side building blocks shared by the sweep and grid kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ephb {

constexpr int kNeighMask = 0x1FFFFFFF;  // LAMMPS NEIGHMASK (reference fix_eph.cpp:452)

// pos4.w carries per-atom bits next to the coordinates so one 32-byte gather
// brings everything a pair evaluation needs about atom j.
constexpr unsigned kElemMask = 0xFFu;    // element index in the .beta file (type_map[type-1])
constexpr unsigned kBitGroup = 1u << 8;  // mask & groupbit
constexpr unsigned kBitValid = 1u << 9;  // rho_j > 0 (reference fix_eph.cpp:768, :811)
constexpr int kTypeShift = 16;           // bits 16..23: LAMMPS type - 1 (the legacy PRB model indexes rho(r) with it)

__device__ __forceinline__ double bits_to_double(unsigned lo) {
  return __hiloint2double(0, static_cast<int>(lo));
}
__device__ __forceinline__ unsigned double_to_bits(double w) { return static_cast<unsigned>(__double2loint(w)); }

// One 256-bit load (LDG.E.256 on sm_100a) of a 32-byte record through the
// read-only path: a gathered record costs one L1TEX request instead of two.
__device__ __forceinline__ double4 ld256(const double4 *p) {
#ifdef EPHA_HOST_EMULATION   // tests/emul: this header compiled for the host (test infrastructure only)
  return *p;
#else
  double4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
#endif
}

// First half of a 32-byte record (LDG.128 through the read-only path).
__device__ __forceinline__ double2 ld128(const double4 *p) {
#ifdef EPHA_HOST_EMULATION
  return double2{p->x, p->y};
#else
  double2 r;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
#endif
}

// Device-side bookkeeping of the two-level Verlet list (see eph_sweeps.cuh).
struct ListState {
  unsigned inner_invalid;   // != 0: the inner list must not be used (some atom moved more than inner_skin/2)
  unsigned pad;
  unsigned long long disp0_sq_bits;  // max squared displacement since LAMMPS built its list (double bits, >= 0)
  double guard_sq;          // the inner list stands while no atom has moved further than this (squared) since its build:
                            // half of the skin the list is COMPLETE for, set by prep_coupling in the step that builds it
};

// Streams that are read or written exactly once per pass (list indices, pair weights): evict-first hints keep
// them from displacing the gathered per-atom records in L1/L2.
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

// Sub-warp groups: LANES consecutive lanes cooperate on one atom.
template <int LANES>
__device__ __forceinline__ constexpr unsigned lanes_bits() {
  return LANES >= 32 ? 0xFFFFFFFFu : ((1u << (LANES & 31)) - 1u);
}
template <int LANES>
__device__ __forceinline__ unsigned group_mask(int lane) {
  return lanes_bits<LANES>() << (lane & ~(LANES - 1) & 31);
}

template <int LANES>
__device__ __forceinline__ double group_sum(double v, unsigned gmask) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// EPH_Spline::operator() (reference eph_spline.h:134-142): truncating index,
// Horner in absolute x.  `tab` holds {a,b} and {c,d} as consecutive double2.
__device__ __forceinline__ double spline_eval(const double2 *__restrict__ tab, double inv_dx, double x) {
  unsigned idx = static_cast<unsigned>(x * inv_dx);
  double2 ab = tab[2 * idx];
  double2 cd = tab[2 * idx + 1];
  return fma(x, fma(x, fma(x, cd.y, cd.x), ab.y), ab.x);
}

// 1/x to full double precision without the slow-path branches of the generic
// division (x is a squared distance: finite, positive, far from the range ends).
__device__ __forceinline__ double fast_rcp(double x) {
#ifdef EPHA_HOST_EMULATION
  return 1.0 / x;
#else
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // rcp.approx.f64 is good to ~2^-23; two Newton steps square that twice.
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
#endif
}

// Philox4x32-10, identical to the definition pinned in oracle/eph_oracle.c.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// xi(seed, step, tag): three standard normals per atom per step, keyed on the
// global atom tag so that a ghost regenerates its owner's numbers locally and
// the reference's XI forward-comm (fix_eph.cpp:863-864) needs no exchange.
__device__ __forceinline__ void xi_stream(unsigned long long seed, unsigned long long step, long long tag, double xi[3]) {
  uint32_t r[4];
  unsigned long long t = static_cast<unsigned long long>(tag);
  philox4x32_10(static_cast<uint32_t>(t), static_cast<uint32_t>(t >> 32), static_cast<uint32_t>(step),
                static_cast<uint32_t>(step >> 32), static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const double two_pi = 6.283185307179586476925286766559;
  const double scale = 1.0 / 4294967296.0;
  double u0 = (r[0] + 0.5) * scale, u1 = (r[1] + 0.5) * scale, u2 = (r[2] + 0.5) * scale, u3 = (r[3] + 0.5) * scale;
  double m = sqrt(-2.0 * log(u0));
  double s, c;
  sincos(two_pi * u1, &s, &c);
  xi[0] = m * c;
  xi[1] = m * s;
  xi[2] = sqrt(-2.0 * log(u2)) * cos(two_pi * u3);
}

// EPH_FDM::get_index (reference eph_fdm.h:494-509): floor, then periodic wrap
// through a second floor so that negative cells fold back into the grid.
struct GridGeom {
  int nx, ny, nz;
  double x0, y0, z0, dx, dy, dz;
};

__device__ __forceinline__ int wrap_cell(double x, double x0, double dx, int n) {
  // floor((x - x0) / dx) as the reference computes it.  The quotient comes from a reciprocal (a quarter of the fp64
  // division's instructions); it can differ from the correctly rounded one by an ulp, which matters only within 1e-9 of a
  // cell boundary -- there the division itself decides.
  const double t = x - x0;
  double q = t * fast_rcp(dx);
  if (fabs(q - rint(q)) < 1e-9 * fmax(1.0, fabs(q))) q = t / dx;
  const int l = static_cast<int>(floor(q));
  if (l >= 0 && l < n) return l;   // inside the grid: the second floor is zero
  // l - floor(double(l) / n) * n of the reference is the non-negative remainder; in integers, without a second
  // fp64 division
  const int m = l % n;
  return m < 0 ? m + n : m;
}

__device__ __forceinline__ int grid_index(const GridGeom &g, double x, double y, double z) {
  int lx = wrap_cell(x, g.x0, g.dx, g.nx);
  int ly = wrap_cell(y, g.y0, g.dy, g.ny);
  int lz = wrap_cell(z, g.z0, g.dz, g.nz);
  return lx + ly * g.nx + lz * g.nx * g.ny;
}

}  // namespace ephb
