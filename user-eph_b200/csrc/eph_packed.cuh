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
// it into a constant bit pattern, and one DADD removes the constant.  Both records keep the low 32 bits of their three
// numbers in words 0, 2, 3 and all high bits in word 1, so a field costs one shift to reach.  The error these records add to rho_i, w_i,
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
constexpr double kQScale = 4398046511104.0;   // 2^42
constexpr unsigned kMagicHi = 0x43380000u;     // high word of 1.5 * 2^52: as_double(kMagicHi : t) = 1.5 * 2^52 + t for t < 2^51

// One 16-byte record, used for positions and for vectors alike:
//   position  w0 = x[31:0], w2 = y[31:0], w3 = z[31:0], w1 = x[41:32] | y[41:32] << 10 | z[41:32] << 20 | flags << 30
//   vector    w0 = m0[31:0], w2 = m1[31:0], w3 = m2[31:0], w1 = m0[39:32] | m1[39:32] << 8 | m2[39:32] << 16 | (E + 127) << 24
//             (m_k = rint(a_k 2^(39 - E)) + 2^39, E the exponent with max |a_k| < 2^E)
struct __align__(16) Words16 { unsigned w0, w1, w2, w3; };
struct __align__(32) Packed32 { Words16 p, b; };
typedef Words16 Block16;

// ---- position ----
__device__ __forceinline__ unsigned long long quantise_coord(double x, double inv_period) {
  double t = x * inv_period;
  t -= floor(t);   // [0, 1)
  return static_cast<unsigned long long>(t * kQScale + 0.5) & kQMask;   // a fraction that rounds up to 1 wraps to 0
}
__device__ __forceinline__ Words16 pack_position(double x, double y, double z, double inv_period, unsigned flags) {
  const unsigned long long qx = quantise_coord(x, inv_period), qy = quantise_coord(y, inv_period), qz = quantise_coord(z, inv_period);
  Words16 p;
  p.w0 = static_cast<unsigned>(qx); p.w2 = static_cast<unsigned>(qy); p.w3 = static_cast<unsigned>(qz);
  p.w1 = static_cast<unsigned>(qx >> 32) | (static_cast<unsigned>(qy >> 32) << 10) | (static_cast<unsigned>(qz >> 32) << 20) | ((flags & 3u) << 30);
  return p;
}
__device__ __forceinline__ unsigned packed_flags(const Words16 &p) { return p.w1 >> 30; }

// The centre atom of a sweep: its coordinates shifted by half a period, so that (q_j - o) mod 2^42 is the displacement
// + 2^41, an unsigned number that needs no sign extension.
struct Centre {
  unsigned long long ox, oy, oz;
};
__device__ __forceinline__ Centre make_centre(const Words16 &p) {
  const unsigned long long qx = p.w0 | (static_cast<unsigned long long>(p.w1 & 0x3FFu) << 32);
  const unsigned long long qy = p.w2 | (static_cast<unsigned long long>((p.w1 >> 10) & 0x3FFu) << 32);
  const unsigned long long qz = p.w3 | (static_cast<unsigned long long>((p.w1 >> 20) & 0x3FFu) << 32);
  Centre c;
  c.ox = (qx - kQHalf) & kQMask;
  c.oy = (qy - kQHalf) & kQMask;
  c.oz = (qz - kQHalf) & kQMask;
  return c;
}
// one coordinate: (hi : lo) - o, reduced modulo 2^42, read as a double, minus 2^41.  Bits of `hi` above the field do not
// matter: borrows travel upwards and the mask removes what is left of them.
__device__ __forceinline__ double coord_delta(unsigned lo, unsigned hi, unsigned long long o) {
  const unsigned long long t = ((static_cast<unsigned long long>(hi) << 32) | lo) - o;
  const unsigned th = (static_cast<unsigned>(t >> 32) & 0x3FFu) | kMagicHi;
  return __hiloint2double(static_cast<int>(th), static_cast<int>(static_cast<unsigned>(t))) -
         __hiloint2double(static_cast<int>(kMagicHi | 0x200u), 0);   // 1.5 * 2^52 + 2^41
}
// displacement j - i in quanta, exact, in [-2^41, 2^41)
__device__ __forceinline__ void displacement(const Centre &c, const Words16 &pj, double &dx, double &dy, double &dz) {
  dx = coord_delta(pj.w0, pj.w1, c.ox);
  dy = coord_delta(pj.w2, pj.w1 >> 10, c.oy);
  dz = coord_delta(pj.w3, pj.w1 >> 20, c.oz);
}

// ---- 3-vector ----
__device__ __forceinline__ Words16 pack_vector(double a, double b, double c, unsigned *status) {
  Words16 r;
  r.w0 = r.w2 = r.w3 = 0u;
  r.w1 = 0x00808080u;   // zero vector: mantissas 0 (biased 2^39), smallest exponent
  const double m = fmax(fabs(a), fmax(fabs(b), fabs(c)));
  const unsigned long long mb = static_cast<unsigned long long>(__double_as_longlong(m));
  const int ef = static_cast<int>((mb >> 52) & 0x7FFull);
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
    r.w0 = static_cast<unsigned>(ia); r.w2 = static_cast<unsigned>(ib); r.w3 = static_cast<unsigned>(ic);
    r.w1 = static_cast<unsigned>((ia >> 32) & 0xFF) | (static_cast<unsigned>((ib >> 32) & 0xFF) << 8) |
           (static_cast<unsigned>((ic >> 32) & 0xFF) << 16) | (static_cast<unsigned>(E + 127) << 24);
    break;
  }
  return r;
}
// byte k of `bytes` into the low byte of `base` (whose own low byte is zero): one PRMT on the device
__device__ __forceinline__ unsigned splice_byte(unsigned bytes, unsigned base, int k) {
#ifdef EPHA_HOST_EMULATION
  return base | ((bytes >> (8 * k)) & 0xFFu);
#else
  return __byte_perm(bytes, base, 0x7650u + static_cast<unsigned>(k));
#endif
}
__device__ __forceinline__ void unpack_vector(const Words16 &r, double &a, double &b, double &c) {
  // exponent field E + 13 + 1023 = (E + 127) + 909, mantissa bit 51 set: the double 2^(E+13) (1.5 + m_biased / 2^52)
  const unsigned base_hi = ((r.w1 >> 4) & 0x0FF00000u) + ((909u << 20) | 0x00080000u);
  const double centre = __hiloint2double(static_cast<int>(base_hi | 0x80u), 0);
  a = __hiloint2double(static_cast<int>(splice_byte(r.w1, base_hi, 0)), static_cast<int>(r.w0)) - centre;
  b = __hiloint2double(static_cast<int>(splice_byte(r.w1, base_hi, 1)), static_cast<int>(r.w2)) - centre;
  c = __hiloint2double(static_cast<int>(splice_byte(r.w1, base_hi, 2)), static_cast<int>(r.w3)) - centre;
}

// ---- record loads through the read-only path ----
__device__ __forceinline__ Packed32 ld_packed32(const Packed32 *p) {
#ifdef EPHA_HOST_EMULATION
  return *p;
#else
  unsigned long long a, b, c, d;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  Packed32 r;
  r.p.w0 = static_cast<unsigned>(a); r.p.w1 = static_cast<unsigned>(a >> 32); r.p.w2 = static_cast<unsigned>(b); r.p.w3 = static_cast<unsigned>(b >> 32);
  r.b.w0 = static_cast<unsigned>(c); r.b.w1 = static_cast<unsigned>(c >> 32); r.b.w2 = static_cast<unsigned>(d); r.b.w3 = static_cast<unsigned>(d >> 32);
  return r;
#endif
}
__device__ __forceinline__ Words16 ld_block16(const Words16 *p) {
#ifdef EPHA_HOST_EMULATION
  return *p;
#else
  Words16 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.w3) : "l"(p));
  return r;
#endif
}

__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}

// Launch shapes, measured on B200 at 4 M atoms (profiles/r2_force_variants.txt): two list slots in flight per lane in both
// passes; density 256 threads x 4 CTAs (63 registers), force 256 threads x 3 CTAs (80 registers).
#ifndef EPH_THREADS_FORCE_PACKED
#define EPH_THREADS_FORCE_PACKED 256
#endif
#ifndef EPH_MINB_FORCE_PACKED
#define EPH_MINB_FORCE_PACKED 3
#endif
#ifndef EPH_FORCE_SLOTS
#define EPH_FORCE_SLOTS 2
#endif
#ifndef EPH_DENSITY_SLOTS
#define EPH_DENSITY_SLOTS 2
#endif
#ifndef EPH_MINB_DENSITY_PACKED
#define EPH_MINB_DENSITY_PACKED 4
#endif
struct PackedArgs {
  const Packed32 *__restrict__ D;   // [ntotal] {position (flags: element index) | v}
  const Packed32 *__restrict__ A;   // [ntotal] {position (flags: bit0 rho > 0, bit1 in group) | u}
  const Words16 *__restrict__ B;    // [ntotal] z
  const double4 *__restrict__ pos4; // [ntotal] fp64 positions + bits of pack_atoms (group bit of the centre atom)
  const double *__restrict__ var;   // [nlocal] eta_factor sqrt(T_e(cell)) of prep_coupling
  double quantum_sq;                // (P / 2^42)^2: squared length of one position quantum
};

// The inner list on its own, for steps that rebuild it and then run the packed sweeps on it: one WARP per atom walks the
// atom's row of LAMMPS' list (136 entries at r_c + 2 A: index loads of 128 bytes, one gathered fp64 position per lane),
// keeps the neighbours closer than r_c + inner_skin in list order (ballot + popc) and writes them into the atom's slots
// of its tile; a warp does the atoms of one tile in turn so that it can pad all of them to the tile's longest list
// (see below).  Exact fp64 positions decide membership, like in density_sweep_kernel<BUILD>.
#ifndef EPH_BUILD_CHUNKS
#define EPH_BUILD_CHUNKS 5
#endif
#ifndef EPH_STAGE_ENTRIES
#define EPH_STAGE_ENTRIES 1536
#endif
constexpr int kBuildChunks = EPH_BUILD_CHUNKS;
constexpr int kStageEntries = EPH_STAGE_ENTRIES;   // list slots of a tile staged in shared memory per warp (6 KB): 96 per atom at 2 lanes

template <int LANES, bool MULTI>
__global__ void __launch_bounds__(256) inner_build_kernel(SweepArgs a, const double4 *__restrict__ pos4) {
  constexpr int TILE = 32 / LANES;
  // The kept neighbours of an atom land 8 bytes per 128-byte line of the tile layout: written straight to global memory
  // they cost as many L1 wavefronts as the gathers.  The warp assembles its tile here and copies it out in whole lines.
  __shared__ int s_stage[8][kStageEntries];
  int *stage = s_stage[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  const int warps_per_block = blockDim.x >> 5;
  const int ntiles = (a.nlocal + TILE - 1) / TILE;
  for (int tile = blockIdx.x * warps_per_block + (threadIdx.x >> 5); tile < ntiles; tile += gridDim.x * warps_per_block) {
    const long long t0 = a.tile_off[tile];
    int cnt_mine = 0, tmax = 0;   // lane t keeps the list length of the tile's atom t
    for (int t = 0; t < TILE; ++t) {
      const int i = tile * TILE + t;
      int cnt = 0;
      if (i < a.nlocal) {
        const double4 pi = pos4[i];
        if (double_to_bits(pi.w) & kBitGroup) {   // atoms outside the group have no list (their rho and w stay zero)
          const long long row = a.offsets[i];
          const int n = static_cast<int>(a.offsets[i + 1] - row);
          // kBuildChunks runs of 32 entries per trip: all their indices, then all their positions are requested before
          // the first one is tested, so a lane has that many gathers in flight (a 136-entry row is one trip)
          for (int k0 = 0; k0 < n; k0 += 32 * kBuildChunks) {
            int j[kBuildChunks];
            bool in[kBuildChunks];
#pragma unroll
            for (int u = 0; u < kBuildChunks; ++u) {
              const int k = k0 + 32 * u + lane;
              j[u] = k < n ? (ld_stream(a.neigh + row + k) & kNeighMask) : -1;
            }
#pragma unroll
            for (int u = 0; u < kBuildChunks; ++u) {
              in[u] = false;
              if (j[u] >= 0) {
                const double4 pj = ld256(pos4 + j[u]);
                const double ex = pj.x - pi.x, ey = pj.y - pi.y, ez = pj.z - pi.z;
                in[u] = ex * ex + ey * ey + ez * ez < a.r_inner_sq;
              }
            }
#pragma unroll
            for (int u = 0; u < kBuildChunks; ++u) {
              if (k0 + 32 * u >= n) break;   // warp-uniform
              const unsigned bal = __ballot_sync(0xFFFFFFFFu, in[u]);
              if (in[u]) {
                const int c = cnt + __popc(bal & below);
                const int pos = (c / LANES) * 32 + t * LANES + (c & (LANES - 1));
                if (pos < kStageEntries) stage[pos] = j[u];
                else a.ineigh[t0 + pos] = j[u];
              }
              cnt += __popc(bal);
            }
          }
        }
        if (lane == 0) a.icount[i] = cnt;
      }
      if (lane == t) cnt_mine = cnt;
      tmax = max(tmax, cnt);
    }
    // padding: every atom's slots up to the tile's longest list, rounded up to kPadIters iterations, hold the atom's own
    // index and a zero pair weight.  Lane l owns column l of the tile: atom l / LANES, every LANES-th list position.
    const int padded = (tmax + kPadIters * LANES - 1) / (kPadIters * LANES) * (kPadIters * LANES);
    const int rows = padded / LANES;                      // warp iterations of the sweeps over this tile
    const int t_mine = lane / LANES;
    const int i_mine = tile * TILE + t_mine;
    const int self = i_mine < a.nlocal ? i_mine : 0;
    const int cnt_col = __shfl_sync(0xFFFFFFFFu, cnt_mine, t_mine);
    __syncwarp();
    for (int r = 0; r < rows; ++r) {
      const int pos = r * 32 + lane;
      const int c = r * LANES + (lane & (LANES - 1));
      const bool pad = c >= cnt_col;
      int v;
      if (pos < kStageEntries) v = pad ? self : stage[pos];
      else v = pad ? self : a.ineigh[t0 + pos];
      a.ineigh[t0 + pos] = v;                              // whole 128-byte lines
      if (pad) {
        st_stream(a.gpair + t0 + pos, 0.0);
        if (MULTI) st_stream(a.gpair_i + t0 + pos, 0.0);
      }
    }
    __syncwarp();
  }
}

// The packed sweeps walk the tiles of the inner list with ONE trip count per warp: the step that builds the list pads
// every atom's slots up to the longest list of its tile (rounded up to two iterations) with the atom's own index and a
// zero pair weight, so the index and weight streams run ahead without bounds tests and every gather of the (warp-wide)
// trip count is legal.
//
// Density pass on packed records.  When the device-side guard has invalidated the inner list the launch returns at once
// and the fp64 kernel that follows does the step on LAMMPS' list.  Two list slots per lane are in flight: both records
// are requested before either is used.
template <int LANES, bool MULTI, bool FRIC>
__global__ void __launch_bounds__(256, EPH_MINB_DENSITY_PACKED) density_packed_kernel(SweepArgs a, PackedArgs q) {
  if (*a.inner_invalid != 0u) return;
  const RhoTable<0> tab{a.rho_tab4, nullptr, nullptr};
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;
  constexpr int TILE = 32 / LANES;
  const int n_work_pad = (a.n_work + TILE - 1) / TILE * TILE;   // whole warps: the trip count is a warp-wide maximum

  for (int w = blockIdx.x * groups_per_block + group_in_block; w < n_work_pad; w += gridDim.x * groups_per_block) {
    const int i = a.work ? a.work[w / TILE] * TILE + (w & (TILE - 1)) : w;
    const bool real = w < a.n_work && i < a.nlocal;
    double rho = 0.0, wx = 0.0, wy = 0.0, wz = 0.0;
    bool active = false;
    Packed32 ri;
    ri.p = Words16{0u, 0u, 0u, 0u}; ri.b = ri.p;
    int nn = 0;
    if (real) {
      ri = ld_packed32(q.D + i);
      active = (double_to_bits(q.pos4[i].w) & kBitGroup) != 0u;   // atoms outside the group: rho = 0, w = 0 (fix_eph.cpp:442-445, :704)
      nn = active ? a.icount[i] : 0;
    }
    constexpr int S = EPH_DENSITY_SLOTS;   // list slots per lane and iteration: all records are requested before any is used
    static_assert(kPadIters % S == 0, "the padding of the tiles covers whole iterations only for divisors of kPadIters");
    const int niter = (warp_max_int(nn) + S * LANES - 1) / (S * LANES);
    const long long first = a.tile_off[(a.work ? i : w) / TILE] + lane;
    const Centre c = make_centre(ri.p);
    const int off_i = MULTI ? (int)packed_flags(ri.p) * a.n_rho : 0;
    double vix = 0.0, viy = 0.0, viz = 0.0;
    if (FRIC) unpack_vector(ri.b, vix, viy, viz);
    const int *__restrict__ lp = a.ineigh + first;
    double *__restrict__ gp = a.gpair + first;
    double *__restrict__ gip = MULTI ? a.gpair_i + first : nullptr;
    // slots k, k + LANES, ...; the tile holds them 32 entries apart
    unsigned jq[S];
#pragma unroll
    for (int t = 0; t < S; ++t) jq[t] = (unsigned)ld_stream(lp + 32 * t);
    for (int m = 0, k = sub; m < niter; ++m, k += S * LANES) {
      Packed32 rq[S];
#pragma unroll
      for (int t = 0; t < S; ++t) rq[t] = ld_packed32(q.D + jq[t]);
#pragma unroll
      for (int t = 0; t < S; ++t) jq[t] = (unsigned)ld_stream(lp + 32 * (S * (m + 1) + t));
#pragma unroll
      for (int h = 0; h < S; ++h) {
        const Packed32 &rj = rq[h];
        if (k + h * LANES >= nn) break;
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
          if (FRIC) {   // fix_eph.cpp:726-738 without the per-atom prefactor alpha_i/rho_i; no test on rho_j
            double vjx, vjy, vjz;
            unpack_vector(rj.b, vjx, vjy, vjz);
            const double d = g * (dx * (vix - vjx) + dy * (viy - vjy) + dz * (viz - vjz));
            wx += d * dx; wy += d * dy; wz += d * dz;
          }
        }
        st_stream(gp + 32 * (S * m + h), g);
        if (MULTI) st_stream(gip + 32 * (S * m + h), gi);
      }
    }
    rho = group_sum<LANES>(rho, gmask);
    if (FRIC) {
      // W accumulated with displacements in quanta: two factors of the quantum bring it to A^2
      wx = group_sum<LANES>(wx, gmask) * q.quantum_sq;
      wy = group_sum<LANES>(wy, gmask) * q.quantum_sq;
      wz = group_sum<LANES>(wz, gmask) * q.quantum_sq;
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
template <int LANES, bool MULTI, bool FRIC, bool RAND>
__global__ void __launch_bounds__(EPH_THREADS_FORCE_PACKED, EPH_MINB_FORCE_PACKED) force_packed_kernel(SweepArgs a, PackedArgs q) {
  if (a.walk_mode == 1 && *a.inner_invalid != 0u) return;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES - 1);
  const unsigned gmask = group_mask<LANES>(lane);
  const int groups_per_block = blockDim.x / LANES;
  const int group_in_block = threadIdx.x / LANES;
  constexpr int TILE = 32 / LANES;
  const int i_end_pad = a.i_begin + (a.i_end - a.i_begin + TILE - 1) / TILE * TILE;   // i_begin is a multiple of 32

  for (int i = a.i_begin + blockIdx.x * groups_per_block + group_in_block; i < i_end_pad; i += gridDim.x * groups_per_block) {
    const bool real = i < a.i_end;
    Packed32 ri;
    ri.p = Words16{0u, 0u, 0u, 0u}; ri.b = ri.p;
    int nn = 0;
    if (real) ri = ld_packed32(q.A + i);
    const bool active = real && packed_flags(ri.p) == 3u;   // in the group and rho_i > 0 (fix_eph.cpp:749-754, :793-798)
    if (active) nn = a.icount[i];
    const int niter = (warp_max_int(nn) + LANES - 1) / LANES;
    const long long first = a.tile_off[i / TILE] + lane;
    double fx = 0, fy = 0, fz = 0, rx = 0, ry = 0, rz = 0;
    const Centre c = make_centre(ri.p);
    double uix = 0, uiy = 0, uiz = 0, zix = 0, ziy = 0, ziz = 0;
    if (FRIC) unpack_vector(ri.b, uix, uiy, uiz);
    if (RAND && active) unpack_vector(ld_block16(q.B + i), zix, ziy, ziz);
    const int *__restrict__ lp = a.ineigh + first;
    const double *__restrict__ gp = a.gpair + first;
    const double *__restrict__ gip = MULTI ? a.gpair_i + first : nullptr;
    // One list slot: both force terms of the pair from its records (nothing happens behind the list, beyond the cut-off --
    // fix_eph.cpp:768, :811 -- or for a neighbour with rho_j <= 0).
    auto pair_terms = [&](int k, double gj, double gi, const Packed32 &rj, const Words16 &bj) {
      if (k >= nn || (gj == 0.0 && gi == 0.0) || !(packed_flags(rj.p) & 1u)) return;
      double dx, dy, dz;
      displacement(c, rj.p, dx, dy, dz);
      if (FRIC) {
        double ux, uy, uz;
        unpack_vector(rj.b, ux, uy, uz);
        const double di = dx * uix + dy * uiy + dz * uiz;
        const double dj = dx * ux + dy * uy + dz * uz;
        const double g = gj * di - gi * dj;
        fx -= g * dx; fy -= g * dy; fz -= g * dz;   // friction is negative, fix_eph.cpp:781-784
      }
      if (RAND) {
        double zx, zy, zz;
        unpack_vector(bj, zx, zy, zz);
        const double di = dx * zix + dy * ziy + dz * ziz;
        const double dj = dx * zx + dy * zy + dz * zz;
        const double g = gj * di - gi * dj;
        rx += g * dx; ry += g * dy; rz += g * dz;   // fix_eph.cpp:823-826
      }
    };
    // EPH_FORCE_SLOTS list slots per lane and iteration, all their records requested before any is evaluated.  Per warp and
    // iteration the pass waits for the slowest of its gathers and then runs a long dependent chain of fp64 work; measured
    // on B200 (profiles/r2_force_variants.txt): one slot in flight 1.53-1.64 ms at 28-32 warps per SM, two slots 1.23 ms
    // at 24 warps (80 registers), four slots 1.32 ms at 16; register-held software pipelines and deeper index queues
    // lose to the occupancy they cost.
    constexpr int S = EPH_FORCE_SLOTS;
    static_assert(kPadIters % S == 0, "the padding of the tiles covers whole iterations only for divisors of kPadIters");
    const int nblk = (niter + S - 1) / S;
    unsigned jq[S];
    double gq[S], giq[S];
#pragma unroll
    for (int t = 0; t < S; ++t) {
      jq[t] = (unsigned)ld_stream(lp + 32 * t);
      gq[t] = ld_stream(gp + 32 * t);
      giq[t] = MULTI ? ld_stream(gip + 32 * t) : 0.0;
    }
#pragma unroll 1
    for (int m = 0; m < nblk; ++m) {
      const int k = sub + S * m * LANES;
      Packed32 ra[S];
      Words16 rb[S];
      double ga[S], gia[S];
#pragma unroll
      for (int t = 0; t < S; ++t) {
        ra[t] = ld_packed32(q.A + jq[t]);
        rb[t] = ri.b;
        if (RAND) rb[t] = ld_block16(q.B + jq[t]);
        ga[t] = gq[t];
        gia[t] = MULTI ? giq[t] : gq[t];
      }
#pragma unroll
      for (int t = 0; t < S; ++t) {
        jq[t] = (unsigned)ld_stream(lp + 32 * (S * (m + 1) + t));
        gq[t] = ld_stream(gp + 32 * (S * (m + 1) + t));
        if (MULTI) giq[t] = ld_stream(gip + 32 * (S * (m + 1) + t));
      }
#pragma unroll
      for (int t = 0; t < S; ++t) pair_terms(k + t * LANES, ga[t], gia[t], ra[t], rb[t]);
    }
    if (FRIC) { fx = group_sum<LANES>(fx, gmask); fy = group_sum<LANES>(fy, gmask); fz = group_sum<LANES>(fz, gmask); }
    if (RAND) { rx = group_sum<LANES>(rx, gmask); ry = group_sum<LANES>(ry, gmask); rz = group_sum<LANES>(rz, gmask); }
    if (real && sub == 0) {
      // displacements were in quanta: two factors of the quantum; the random force also takes
      // eta_factor sqrt(T_e(nearest cell)) (fix_eph.cpp:829-833), worked out once per atom by prep_coupling
      double var = 0.0;
      if (RAND && active) var = q.var[i] * q.quantum_sq;
      fx *= q.quantum_sq; fy *= q.quantum_sq; fz *= q.quantum_sq;
      rx *= var; ry *= var; rz *= var;
      const size_t o = 3 * (size_t)i;
      if (FRIC) { a.f_eph[o] = fx; a.f_eph[o + 1] = fy; a.f_eph[o + 2] = fz; }
      if (RAND) { a.f_rng[o] = rx; a.f_rng[o + 1] = ry; a.f_rng[o + 2] = rz; }
      // f += f_EPH (+ f_RNG) for every local atom, grouped or not (fix_eph.cpp:892-906)
      if (a.f != nullptr) {
        double ax = 0, ay = 0, az = 0;
        if (FRIC && a.add_friction) { ax += fx; ay += fy; az += fz; }
        if (RAND && a.add_random) { ax += rx; ay += ry; az += rz; }
        a.f[o] += ax; a.f[o + 1] += ay; a.f[o + 2] += az;
      }
    }
  }
}

}  // namespace ephb
