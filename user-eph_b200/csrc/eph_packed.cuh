// Packed gather records and the two list sweeps that walk them (sm_100a).
//
// The sweeps of eph_sweeps.cuh are bound by L1TEX wavefronts: one per 32-byte sector a gather touches
// (profiles/r1_final_4M_ncu_summary.csv: data pipe at 97 % / 88 %, DRAM at 20 %).  At fp64 a pair costs two sectors in
// the density pass ({x,y,z,bits | v}) and three in the force pass ({x,y,z,bits | u,z}).  The records here carry the
// same information in half the sectors:
//
//   position  16 bytes   three 42-bit fixed-point fractions of a period P (a power of two > 2 (r_c + 2 inner_skin),
//                        16 A at the defaults) plus two flag bits.  Differences are taken modulo 2^42, so the minimum
//                        image inside the period comes out of the integer subtraction itself; both atoms of a pair are
//                        quantised the same way, e_ij = -e_ji holds exactly and momentum conservation survives.
//                        Quantum P / 2^42 = 3.6e-12 A at P = 16 A (error of a displacement <= one quantum).
//   3-vector  16 bytes   block floating point: three 40-bit mantissas and one shared 8-bit exponent (v, u = s w,
//                        z = s xi).  Error <= 2^-39 of the largest component (1.8e-12).
//
//   D[a] = {position | v}                  density pass:  1 sector per list slot (was 2)
//   A[a] = {position + valid | u}, B[a] = {z}  force pass: 2 sectors per slot (was 3), 1 without the random force
//
// Both decodes avoid integer <-> fp64 conversions: a 42-bit integer t is read as the double 1.5 * 2^52 + t by OR-ing
// it into a constant bit pattern, and one DADD removes the constant.  The error these records add to rho_i, w_i,
// f_EPH and f_RNG is bounded in DESIGN.md section 4 (<= ~1e-11 of the largest value, against the 1e-10 bar) and
// measured in every parity test; EPH_B200_RECORDS=exact (or eph_b200_set_precision) selects the fp64 records instead.
//
// The packed sweeps only ever walk the INNER list (pairs closer than r_c + 2 inner_skin by the device-side
// displacement guard), so |e| < P / 2 holds by construction and the modular difference cannot alias; steps that walk
// LAMMPS' list (inner-list builds, guard trips) run the fp64 kernels of eph_sweeps.cuh.
#pragma once

#include "eph_sweeps.cuh"

namespace ephb {

constexpr unsigned long long kQMask = (1ull << 42) - 1ull;
constexpr unsigned long long kQHalf = 1ull << 41;
constexpr double kQScale = 4398046511104.0;                 // 2^42
constexpr unsigned long long kMagic = 0x4338000000000000ull;   // bit pattern of 1.5 * 2^52

struct __align__(16) Packed16 { unsigned long long w0, w1; };
struct __align__(16) Block16 { unsigned w0, w1, w2, w3; };
struct __align__(32) Packed32 { Packed16 p; Block16 b; };

// ---- position: x | y | z as 42-bit fractions of the period ----
//   w0 = x (bits 0..41) | y[41:20] (bits 42..63)
//   w1 = flags (bits 0..1) | y[19:0] (bits 2..21) | z (bits 22..63)
__device__ __forceinline__ unsigned long long quantise_coord(double x, double inv_period) {
  double t = x * inv_period;
  t -= floor(t);   // [0, 1)
  return static_cast<unsigned long long>(t * kQScale + 0.5) & kQMask;   // a fraction that rounds up to 1 wraps to 0
}
__device__ __forceinline__ Packed16 pack_position(double x, double y, double z, double inv_period, unsigned flags) {
  const unsigned long long qx = quantise_coord(x, inv_period), qy = quantise_coord(y, inv_period), qz = quantise_coord(z, inv_period);
  Packed16 p;
  p.w0 = qx | ((qy >> 20) << 42);
  p.w1 = (unsigned long long)(flags & 3u) | ((qy & 0xFFFFFull) << 2) | (qz << 22);
  return p;
}
__device__ __forceinline__ unsigned packed_flags(const Packed16 &p) { return static_cast<unsigned>(p.w1) & 3u; }

// The centre atom of a sweep: its coordinates shifted by half a period, so that (q_j - o) mod 2^42 is the unsigned
// displacement + 2^41 and the signed displacement needs no sign extension.
struct Centre {
  unsigned long long ox, oy, ozs;   // ozs: o_z << 22 (z sits in the top 42 bits of w1)
};
__device__ __forceinline__ Centre make_centre(const Packed16 &p) {
  const unsigned long long qx = p.w0 & kQMask;
  const unsigned long long qy = ((p.w0 >> 42) << 20) | ((p.w1 >> 2) & 0xFFFFFull);
  const unsigned long long qz = p.w1 >> 22;
  Centre c;
  c.ox = (qx - kQHalf) & kQMask;
  c.oy = (qy - kQHalf) & kQMask;
  c.ozs = ((qz - kQHalf) & kQMask) << 22;
  return c;
}
// displacement j - i in quanta, exact, in [-2^41, 2^41)
__device__ __forceinline__ void displacement(const Centre &c, const Packed16 &pj, double &dx, double &dy, double &dz) {
  const double half = __longlong_as_double(static_cast<long long>(kMagic | kQHalf));
  const unsigned long long tx = (pj.w0 - c.ox) & kQMask;
  const unsigned long long qy = ((pj.w0 >> 42) << 20) | ((pj.w1 >> 2) & 0xFFFFFull);
  const unsigned long long ty = (qy - c.oy) & kQMask;
  const unsigned long long tz = (pj.w1 - c.ozs) >> 22;   // the low 22 bits of ozs are zero: no borrow out of the flag / y bits
  dx = __longlong_as_double(static_cast<long long>(kMagic | tx)) - half;
  dy = __longlong_as_double(static_cast<long long>(kMagic | ty)) - half;
  dz = __longlong_as_double(static_cast<long long>(kMagic | tz)) - half;
}

// ---- 3-vector: mantissas m_k = rint(a_k 2^(39 - E)) stored biased by 2^39, E the exponent with max |a_k| < 2^E ----
//   w0, w1, w2 = low 32 bits of the three biased mantissas; w3 = their high bytes (bytes 0..2) | (E + 127) << 24
__device__ __forceinline__ Block16 pack_vector(double a, double b, double c, unsigned *status) {
  Block16 r;
  r.w0 = r.w1 = r.w2 = 0u;
  r.w3 = 0x00808080u;   // zero vector: mantissas 0 (biased 2^39), smallest exponent
  const double m = fmax(fabs(a), fmax(fabs(b), fabs(c)));
  const unsigned long long mb = static_cast<unsigned long long>(__double_as_longlong(m));
  int ef = static_cast<int>((mb >> 52) & 0x7FFull);
  if (ef == 0x7FF || a != a || b != b || c != c) {   // non-finite input: flagged, encoded as zero
    if (status) atomicOr(status, 4u);
    return r;
  }
  if (ef == 0) return r;                      // zero / subnormal
  int E = ef - 1022;                          // m = f 2^E, f in [0.5, 1)
  if (E < -127) return r;                     // below 2^-128: indistinguishable from zero at this scale
  if (E > 128) { if (status) atomicOr(status, 4u); E = 128; }
  for (int pass = 0; pass < 2; ++pass) {
    const double scale = __longlong_as_double(static_cast<long long>(static_cast<unsigned long long>(1023 + 39 - E) << 52));
    const double ma = rint(a * scale), mbv = rint(b * scale), mc = rint(c * scale);
    const double top = 549755813888.0;   // 2^39
    if (pass == 0 && E < 128 && (fabs(ma) >= top || fabs(mbv) >= top || fabs(mc) >= top)) { ++E; continue; }
    const double lim = top - 1.0;
    const long long ia = static_cast<long long>(fmin(fmax(ma, -lim), lim)) + (1ll << 39);
    const long long ib = static_cast<long long>(fmin(fmax(mbv, -lim), lim)) + (1ll << 39);
    const long long ic = static_cast<long long>(fmin(fmax(mc, -lim), lim)) + (1ll << 39);
    r.w0 = static_cast<unsigned>(ia); r.w1 = static_cast<unsigned>(ib); r.w2 = static_cast<unsigned>(ic);
    r.w3 = static_cast<unsigned>((ia >> 32) & 0xFF) | (static_cast<unsigned>((ib >> 32) & 0xFF) << 8) |
           (static_cast<unsigned>((ic >> 32) & 0xFF) << 16) | (static_cast<unsigned>(E + 127) << 24);
    break;
  }
  return r;
}
__device__ __forceinline__ void unpack_vector(const Block16 &r, double &a, double &b, double &c) {
  // exponent field E + 13 + 1023 = (E + 127) + 909, mantissa bit 51 set: the double 2^(E+13) (1.5 + m_biased / 2^52)
  const unsigned base_hi = ((r.w3 >> 4) & 0x0FF00000u) + ((909u << 20) | 0x00080000u);
  const double centre = __hiloint2double(static_cast<int>(base_hi | 0x80u), 0);
  a = __hiloint2double(static_cast<int>(base_hi | (r.w3 & 0xFFu)), static_cast<int>(r.w0)) - centre;
  b = __hiloint2double(static_cast<int>(base_hi | ((r.w3 >> 8) & 0xFFu)), static_cast<int>(r.w1)) - centre;
  c = __hiloint2double(static_cast<int>(base_hi | ((r.w3 >> 16) & 0xFFu)), static_cast<int>(r.w2)) - centre;
}

// ---- record loads through the read-only path ----
__device__ __forceinline__ Packed32 ld_packed32(const Packed32 *p) {
#ifdef EPHA_HOST_EMULATION
  return *p;
#else
  Packed32 r;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.p.w0), "=l"(r.p.w1), "=l"(*reinterpret_cast<unsigned long long *>(&r.b.w0)),
                 "=l"(*reinterpret_cast<unsigned long long *>(&r.b.w2))
               : "l"(p));
  return r;
#endif
}
__device__ __forceinline__ Block16 ld_block16(const Block16 *p) {
#ifdef EPHA_HOST_EMULATION
  return *p;
#else
  Block16 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.w3) : "l"(p));
  return r;
#endif
}

struct PackedArgs {
  const Packed32 *__restrict__ D;   // [ntotal] {position (flags: element index) | v}
  const Packed32 *__restrict__ A;   // [ntotal] {position (flags: bit0 rho > 0, bit1 in group) | u}
  const Block16 *__restrict__ B;    // [ntotal] z
  const double4 *__restrict__ pos4; // [ntotal] fp64 positions + bits of pack_atoms (group bit of the centre atom)
  const double *__restrict__ var;   // [nlocal] eta_factor sqrt(T_e(cell)) of prep_coupling
  double quantum_sq;                // (P / 2^42)^2: squared length of one position quantum
};

// Density pass on packed records.  Walks the inner list only; when the device-side guard has invalidated it the
// launch returns at once and the fp64 kernel that follows does the step on LAMMPS' list.
// Two list slots per lane are in flight: both records are requested before either is used.
template <int LANES, bool MULTI>
__global__ void __launch_bounds__(256, EPH_MINB_DENSITY) density_packed_kernel(SweepArgs a, PackedArgs q) {
  if (*a.inner_invalid != 0u) return;
  const RhoTable<0> tab{a.rho_tab4, nullptr, nullptr};
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;

  for (int w = blockIdx.x * groups_per_block + group_in_block; w < a.n_work; w += gridDim.x * groups_per_block) {
    const int i = a.work ? a.work[w / (32 / LANES)] * (32 / LANES) + (w & (32 / LANES - 1)) : w;
    const bool real = i < a.nlocal;
    double rho = 0.0, wx = 0.0, wy = 0.0, wz = 0.0;
    bool active = false;
    Packed32 ri;
    int nn = 0;
    long long first = 0;
    if (real) {
      ri = ld_packed32(q.D + i);
      active = (double_to_bits(q.pos4[i].w) & kBitGroup) != 0u;   // atoms outside the group: rho = 0, w = 0 (fix_eph.cpp:442-445, :704)
      nn = a.icount[i];
      first = a.tile_off[i / (32 / LANES)] + lane;
    }
    if (active) {
      const Centre c = make_centre(ri.p);
      const int off_i = MULTI ? (int)packed_flags(ri.p) * a.n_rho : 0;
      double vix = 0.0, viy = 0.0, viz = 0.0;
      if (a.do_friction) unpack_vector(ri.b, vix, viy, viz);
      const int *__restrict__ lp = a.ineigh + first;
      double *__restrict__ gp = a.gpair + first;
      double *__restrict__ gip = MULTI ? a.gpair_i + first : nullptr;
      // slots k (this lane) and k + LANES; the tile holds them 32 entries apart
      int ja = sub < nn ? ld_stream(lp) : i;
      int jb = sub + LANES < nn ? ld_stream(lp + 32) : i;
      for (int k = sub, slot = 0; k < nn; k += 2 * LANES, slot += 64) {
        const bool have_b = k + LANES < nn;
        const Packed32 ra = ld_packed32(q.D + (ja & kNeighMask));
        const Packed32 rb = ld_packed32(q.D + (jb & kNeighMask));
        if (k + 2 * LANES < nn) ja = ld_stream(lp + slot + 64);
        if (k + 3 * LANES < nn) jb = ld_stream(lp + slot + 96);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const Packed32 &rj = h ? rb : ra;
          if (h && !have_b) break;
          double dx, dy, dz;
          displacement(c, rj.p, dx, dy, dz);
          const double r2 = (dx * dx + dy * dy + dz * dz) * q.quantum_sq;
          double g = 0.0, gi = 0.0;
          if (r2 < a.r_cutoff_sq) {   // strict '<' as in fix_eph.cpp:457, :724
            const int off_j = MULTI ? (int)packed_flags(rj.p) * a.n_rho : 0;
            const double rho_j = tab.eval(off_j, a.inv_dr_sq, r2);
            const double rinv = fast_rcp(r2);
            g = rho_j * rinv;
            rho += rho_j;
            if (MULTI) gi = (off_j == off_i) ? g : tab.eval(off_i, a.inv_dr_sq, r2) * rinv;
            if (a.do_friction) {   // fix_eph.cpp:726-738 without the per-atom prefactor alpha_i/rho_i; no test on rho_j
              double vjx, vjy, vjz;
              unpack_vector(rj.b, vjx, vjy, vjz);
              const double d = g * (dx * (vix - vjx) + dy * (viy - vjy) + dz * (viz - vjz));
              wx += d * dx; wy += d * dy; wz += d * dz;
            }
          }
          st_stream(gp + slot + 32 * h, g);
          if (MULTI) st_stream(gip + slot + 32 * h, gi);
        }
      }
      rho = group_sum<LANES>(rho, gmask);
      if (a.do_friction) {
        // W accumulated with displacements in quanta: two factors of the quantum bring it to A^2
        wx = group_sum<LANES>(wx, gmask) * q.quantum_sq;
        wy = group_sum<LANES>(wy, gmask) * q.quantum_sq;
        wz = group_sum<LANES>(wz, gmask) * q.quantum_sq;
      }
    }
    if (real && sub == 0) {
      a.rho[i] = rho;
      a.W4[i] = make_double4(wx, wy, wz, 0.0);
    }
    if (a.done_counter != nullptr && w < a.n_boundary) {   // see density_sweep_kernel
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(a.done_counter, 1u);
      }
    }
  }
}

// Force pass on packed records (inner list; the caller launches it only when this step's pair weights are stored in
// the inner list's tiles -- walk_mode 2, or 1 with the guard intact, which is re-checked here).
template <int LANES, bool MULTI>
__global__ void __launch_bounds__(EPH_THREADS_FORCE, EPH_MINB_FORCE) force_packed_kernel(SweepArgs a, PackedArgs q) {
  if (a.walk_mode == 1 && *a.inner_invalid != 0u) return;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;

  for (int i = a.i_begin + blockIdx.x * groups_per_block + group_in_block; i < a.i_end; i += gridDim.x * groups_per_block) {
    const Packed32 ri = ld_packed32(q.A + i);
    const int nn = a.icount[i];
    const long long first = a.tile_off[i / (32 / LANES)] + lane;
    double fx = 0, fy = 0, fz = 0, rx = 0, ry = 0, rz = 0;
    const bool active = packed_flags(ri.p) == 3u;   // in the group and rho_i > 0 (fix_eph.cpp:749-754, :793-798)
    if (active) {
      const Centre c = make_centre(ri.p);
      double uix = 0, uiy = 0, uiz = 0, zix = 0, ziy = 0, ziz = 0;
      if (a.do_friction) unpack_vector(ri.b, uix, uiy, uiz);
      if (a.do_random) unpack_vector(ld_block16(q.B + i), zix, ziy, ziz);
      const int *__restrict__ lp = a.ineigh + first;
      const double *__restrict__ gp = a.gpair + first;
      const double *__restrict__ gip = MULTI ? a.gpair_i + first : nullptr;
      int jn = 0;
      double gjn = 0.0, gin = 0.0;
      if (sub < nn) {
        jn = ld_stream(lp);
        gjn = ld_stream(gp);
        if (MULTI) gin = ld_stream(gip);
      }
#pragma unroll 1
      for (int k = sub, slot = 0; k < nn; k += LANES, slot += 32) {
        const int j = jn & kNeighMask;
        const double gj = gjn;
        const double gi = MULTI ? gin : gjn;
        if (k + LANES < nn) {
          jn = ld_stream(lp + slot + 32);
          gjn = ld_stream(gp + slot + 32);
          if (MULTI) gin = ld_stream(gip + slot + 32);
        }
        if (gj == 0.0 && gi == 0.0) continue;   // beyond the cut-off (fix_eph.cpp:768, :811) or a vanishing pair weight
        const Packed32 rj = ld_packed32(q.A + j);
        Block16 bj;
        if (a.do_random) bj = ld_block16(q.B + j);
        if (!(packed_flags(rj.p) & 1u)) continue;   // rho_j > 0 required, fix_eph.cpp:768, :811
        double dx, dy, dz;
        displacement(c, rj.p, dx, dy, dz);
        if (a.do_friction) {
          double ux, uy, uz;
          unpack_vector(rj.b, ux, uy, uz);
          const double di = dx * uix + dy * uiy + dz * uiz;
          const double dj = dx * ux + dy * uy + dz * uz;
          const double g = gj * di - gi * dj;
          fx -= g * dx; fy -= g * dy; fz -= g * dz;   // friction is negative, fix_eph.cpp:781-784
        }
        if (a.do_random) {
          double zx, zy, zz;
          unpack_vector(bj, zx, zy, zz);
          const double di = dx * zix + dy * ziy + dz * ziz;
          const double dj = dx * zx + dy * zy + dz * zz;
          const double g = gj * di - gi * dj;
          rx += g * dx; ry += g * dy; rz += g * dz;   // fix_eph.cpp:823-826
        }
      }
      fx = group_sum<LANES>(fx, gmask); fy = group_sum<LANES>(fy, gmask); fz = group_sum<LANES>(fz, gmask);
      rx = group_sum<LANES>(rx, gmask); ry = group_sum<LANES>(ry, gmask); rz = group_sum<LANES>(rz, gmask);
    }
    if (sub == 0) {
      // displacements were in quanta: two factors of the quantum; the random force also takes
      // eta_factor sqrt(T_e(nearest cell)) (fix_eph.cpp:829-833), worked out once per atom by prep_coupling
      double var = 0.0;
      if (active && a.do_random) var = q.var[i] * q.quantum_sq;
      fx *= q.quantum_sq; fy *= q.quantum_sq; fz *= q.quantum_sq;
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
