// TMA-tiled form of the FDM sub-step (reference eph_fdm.h:319-395) for sm_100a.
//
// One CTA owns a TX x TY x TZ tile of cells.  The T_e and kappa_e boxes with a one-cell halo are brought into shared
// memory by the Tensor Memory Accelerator (cp.async.bulk.tensor.3d, one elected thread, completion on an mbarrier);
// the TMA zero-fills whatever lies outside the grid, and the threads of tiles that touch the periodic boundary patch
// those halo cells with the wrapped values.  All stencil neighbours are then read from shared memory, the five
// single-use fields (dT_e, S_e, rho_e, C_e, flags) stream in with coalesced loads, T_e streams out.
// Algorithmic traffic: 60 bytes per cell-update; the halo adds (36*10*10)/(32*8*8) - 1 = 76 % to the
// T_e and kappa_e reads, which L2 absorbs.
#pragma once

#include <cuda.h>

#include "eph_grid.cuh"

namespace ephb {

constexpr int kTX = 32, kTY = 8, kTZ = 8;
// fp64 boxes must start at an even x coordinate (16-byte aligned start; an odd start raises `illegal instruction`,
// tools/microbench/tma_probe.cu), so the box carries a two-cell margin in x and a one-cell halo in y and z
constexpr int kHX = 2;
constexpr int kBX = kTX + 2 * kHX, kBY = kTY + 2, kBZ = kTZ + 2;
constexpr int kBoxCells = kBX * kBY * kBZ;
constexpr int kBoxBytes = (kBoxCells * 8 + 127) / 128 * 128;
constexpr int kFlagBytes = (kBoxCells * 2 + 15) / 16 * 16;
constexpr int kTmaSmemBytes = 2 * kBoxBytes + kFlagBytes + 16;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// The CTA waits for a TMA transaction with ONE polling warp (256 threads spinning on try_wait cost the
// constant-coefficient kernel a large share of its issue slots, ncu); the other warps sleep in __syncthreads and then
// observe the completed phase themselves with a single non-blocking test_wait, which is their own acquire of the
// data the TMA wrote.
__device__ __forceinline__ void cta_wait_tma(unsigned long long *bar, unsigned parity, int tid) {
  if (tid < 32) {
    unsigned done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
  }
  __syncthreads();
  asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1; }"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct GridTmaArgs {
  GridArgs g;
  int has_walls;   // some cell has flag == 2: the wall substitution needs the neighbours' flags
};

// The single-use fields of kBatch z-planes of one thread's (i, j) column: fetched together (more bytes in flight per
// thread), and one batch ahead of the stencil so that their latency overlaps the TMA's and the previous batch's math.
constexpr int kBatch = 2;
struct PlaneBatch {
  double dT[kBatch], S[kBatch], rho[kBatch], C[kBatch];
  int fl[kBatch], td[kBatch];
};

__device__ __forceinline__ void load_batch(PlaneBatch &b, const GridArgs &g, int i, int j, int k0, long long sy, long long sz) {
#pragma unroll
  for (int q = 0; q < kBatch; ++q) {
    const long long r = i + j * sy + (long long)min(k0 + q, g.nz - 1) * sz;
    b.fl[q] = g.flag[r];
    b.dT[q] = g.dT_e[r]; b.S[q] = g.S_e[r]; b.rho[q] = g.rho_e[r]; b.C[q] = g.C_e[r];
    b.td[q] = g.E_e_T != nullptr ? g.t_dyn[r] : 0;
  }
}

__global__ void __launch_bounds__(256, 3) fdm_substep_tma_kernel(const __grid_constant__ CUtensorMap map_T,
                                                              const __grid_constant__ CUtensorMap map_K, GridTmaArgs ta) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  double *sT = reinterpret_cast<double *>(tma_smem);                 // kBoxBytes each, 128-byte aligned
  double *sK = reinterpret_cast<double *>(tma_smem + kBoxBytes);
  short *sF = reinterpret_cast<short *>(tma_smem + 2 * kBoxBytes);
  unsigned long long &bar = *reinterpret_cast<unsigned long long *>(tma_smem + 2 * kBoxBytes + kFlagBytes);
  const GridArgs &g = ta.g;
  const int ox = blockIdx.x * kTX, oy = blockIdx.y * kTY, oz = g.z_begin + blockIdx.z * kTZ;
  const int tid = threadIdx.x;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)),
                 "r"(static_cast<unsigned>(2 * kBoxCells * sizeof(double))) : "memory");
    tma_load_3d(sT, &map_T, ox - kHX, oy - 1, oz - 1, &bar);
    tma_load_3d(sK, &map_K, ox - kHX, oy - 1, oz - 1, &bar);
  }
  const long long sy = g.nx, sz = (long long)g.nx * g.ny;
  // flags of the box (2 bytes per cell, periodic wrap applied directly) while the TMA is in flight
  if (ta.has_walls) {
    for (int c = tid; c < kBoxCells; c += blockDim.x) {
      const int bx = c % kBX, by = (c / kBX) % kBY, bz = c / (kBX * kBY);
      int gx = ox - kHX + bx, gy = oy - 1 + by, gz = oz - 1 + bz;
      gx = gx < 0 ? gx + g.nx : (gx >= g.nx ? gx - g.nx : gx);
      gy = gy < 0 ? gy + g.ny : (gy >= g.ny ? gy - g.ny : gy);
      gz = gz < 0 ? gz + g.nz : (gz >= g.nz ? gz - g.nz : gz);
      const bool ok = gx >= 0 && gx < g.nx && gy >= 0 && gy < g.ny && gz >= 0 && gz < g.nz;
      sF[c] = ok ? g.flag[gx + gy * sy + gz * sz] : (short)1;
    }
  }
  const int tx = tid & 31, ty = tid >> 5;   // 32 x 8 threads, each marching over the tile's TZ planes
  const int i = ox + tx, j = oy + ty;
  const bool inside = i < g.nx && j < g.ny;
  PlaneBatch cur;
  if (inside) load_batch(cur, g, i, j, oz, sy, sz);   // in flight together with the two TMA boxes
  cta_wait_tma(&bar, 0u, tid);   // both boxes have landed
  // periodic wrap: halo cells outside the grid were zero-filled by the TMA
  const bool edge = ox == 0 || oy == 0 || oz == 0 || ox + kTX >= g.nx || oy + kTY >= g.ny || oz + kTZ >= g.nz;
  if (edge) {
    for (int c = tid; c < kBoxCells; c += blockDim.x) {
      const int bx = c % kBX, by = (c / kBX) % kBY, bz = c / (kBX * kBY);
      int gx = ox - kHX + bx, gy = oy - 1 + by, gz = oz - 1 + bz;
      if (gx >= 0 && gx < g.nx && gy >= 0 && gy < g.ny && gz >= 0 && gz < g.nz) continue;
      gx = gx < 0 ? gx + g.nx : (gx >= g.nx ? gx - g.nx : gx);
      gy = gy < 0 ? gy + g.ny : (gy >= g.ny ? gy - g.ny : gy);
      gz = gz < 0 ? gz + g.nz : (gz >= g.nz ? gz - g.nz : gz);
      if (gx < 0 || gx >= g.nx || gy < 0 || gy >= g.ny || gz < 0 || gz >= g.nz) continue;  // beyond one period: unused cells
      const long long r = gx + gy * sy + gz * sz;
      sT[c] = g.T_in[r];
      sK[c] = g.kappa_e[r];
    }
  }
  __syncthreads();

  if (!inside) return;
#pragma unroll
  for (int t0 = 0; t0 < kTZ; t0 += kBatch) {
    PlaneBatch nxt;
    if (t0 + kBatch < kTZ) load_batch(nxt, g, i, j, oz + t0 + kBatch, sy, sz);   // next batch ahead of this batch's math
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int tz = t0 + b, k = oz + tz;
      if (k >= g.z_end) break;
      const long long r = i + j * sy + k * sz;
      const int c = (tx + kHX) + (ty + 1) * kBX + (tz + 1) * kBX * kBY;
      double T = sT[c];
      if (cur.fl[b] == 1) {  // only DYNAMIC cells change; walls (2) and constant cells (0) keep T
        const double kr = sK[c];
        double ddT = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int stride = d == 0 ? 1 : (d == 1 ? kBX : kBX * kBY);
          const double inv = d == 0 ? g.inv_dx2 : (d == 1 ? g.inv_dy2 : g.inv_dz2);
          int p = c - stride, q = c + stride;
          if (ta.has_walls) {  // zero-derivative wall substitution, eph_fdm.h:336-337
            if (sF[q] == 2) q = c; else if (sF[p] == 2) p = c;
          }
          const double Tq = sT[q], Tp = sT[p];
          ddT += (sK[q] - sK[p]) * (Tq - Tp) * inv * 0.25;
          ddT += kr * ((Tq + Tp - 2.0 * T) * inv);
        }
        const double src = ddT + cur.dT[b] + cur.S[b];
        if (cur.td[b] == 1) {
          double E = linear_eval(g.E_e_T, g.n_T, g.dT, T);
          E += src / cur.rho[b] * g.inner_dt;
          T = linear_reverse(g.E_e_T, g.n_T, g.dT, E);
        } else {
          T += src / (cur.rho[b] * cur.C[b]) * g.inner_dt;
        }
      }
      if (T < 0.0) {
        T = 0.0;
        atomicOr(g.status, 2u);
      }
      g.T_out[r] = T;
      if (g.clear_source) g.dT_e[r] = 0.0;
    }
    if (t0 + kBatch < kTZ) cur = nxt;
  }
}

// Constant-coefficient fast path: every cell DYNAMIC with the same rho_e, C_e, kappa_e and S_e (what `fix eph ... NX NY NZ
// NULL ...` creates, reference eph_fdm.h:28-46, :143-153).  kappa differences vanish identically, the parameters are
// scalars, and a cell-update moves 24 bytes (T_e in/out, dT_e in) instead of 60.  Both the T_e box (with halo) and the
// dT_e tile arrive by TMA; arithmetic order is that of the general kernel.  The one division per cell has a constant
// denominator: it is done as q = x * (1/d), q += fma(-d, q, x) * (1/d) (correctly rounded but for rare ties; the
// generic fp64 division subroutine was 35 % of this issue-bound kernel's instructions, ncu).
struct GridUniformArgs {
  int nx, ny, nz;
  const double *__restrict__ T_in;   // for the periodic halo patch
  double *__restrict__ T_out;
  double *__restrict__ dT_e;
  double kappa, S, rho, C;
  double inv_rho_C;                  // 1 / (rho * C), correctly rounded on the host
  double inv_dx2, inv_dy2, inv_dz2, inner_dt;
  int clear_source;
  unsigned *__restrict__ status;
  int z_begin, z_end;                // planes updated by this launch (GridArgs)
};

constexpr int kTileBytes = kTX * kTY * kTZ * 8;
constexpr int kUniSmemBytes = kBoxBytes + kTileBytes + 16;

__global__ void __launch_bounds__(256) fdm_uniform_tma_kernel(const __grid_constant__ CUtensorMap map_T,
                                                              const __grid_constant__ CUtensorMap map_S, GridUniformArgs g) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  double *sT = reinterpret_cast<double *>(tma_smem);
  double *sS = reinterpret_cast<double *>(tma_smem + kBoxBytes);
  unsigned long long &bar = *reinterpret_cast<unsigned long long *>(tma_smem + kBoxBytes + kTileBytes);
  const int ox = blockIdx.x * kTX, oy = blockIdx.y * kTY, oz = g.z_begin + blockIdx.z * kTZ;
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)),
                 "r"(static_cast<unsigned>(kBoxCells * 8 + kTileBytes)) : "memory");
    tma_load_3d(sT, &map_T, ox - kHX, oy - 1, oz - 1, &bar);
    tma_load_3d(sS, &map_S, ox, oy, oz, &bar);
  }
  cta_wait_tma(&bar, 0u, tid);
  const long long sy = g.nx, sz = (long long)g.nx * g.ny;
  const bool edge = ox == 0 || oy == 0 || oz == 0 || ox + kTX >= g.nx || oy + kTY >= g.ny || oz + kTZ >= g.nz;
  if (edge) {
    for (int c = tid; c < kBoxCells; c += blockDim.x) {
      const int bx = c % kBX, by = (c / kBX) % kBY, bz = c / (kBX * kBY);
      int gx = ox - kHX + bx, gy = oy - 1 + by, gz = oz - 1 + bz;
      if (gx >= 0 && gx < g.nx && gy >= 0 && gy < g.ny && gz >= 0 && gz < g.nz) continue;
      gx = gx < 0 ? gx + g.nx : (gx >= g.nx ? gx - g.nx : gx);
      gy = gy < 0 ? gy + g.ny : (gy >= g.ny ? gy - g.ny : gy);
      gz = gz < 0 ? gz + g.nz : (gz >= g.nz ? gz - g.nz : gz);
      if (gx < 0 || gx >= g.nx || gy < 0 || gy >= g.ny || gz < 0 || gz >= g.nz) continue;
      sT[c] = g.T_in[gx + gy * sy + gz * sz];
    }
    __syncthreads();
  }
  const int tx = tid & 31, ty = tid >> 5;
  const int i = ox + tx, j = oy + ty;
  if (i >= g.nx || j >= g.ny) return;
  const double prescaler = g.rho * g.C, inv = g.inv_rho_C;
  int c = (tx + kHX) + (ty + 1) * kBX + kBX * kBY;
  double Tm = sT[c - kBX * kBY], T = sT[c];
  const long long r0 = i + j * sy + (long long)oz * sz;
  double *__restrict__ out = g.T_out + r0;
  double *__restrict__ src = g.dT_e + r0;
  const double *__restrict__ sSp = sS + tx + ty * kTX;
#pragma unroll
  for (int tz = 0; tz < kTZ; ++tz, c += kBX * kBY) {
    const double Tn = sT[c + kBX * kBY];
    if (oz + tz < g.z_end) {
      double ddT = 0.0;
      ddT += g.kappa * ((sT[c + 1] + sT[c - 1] - 2.0 * T) * g.inv_dx2);
      ddT += g.kappa * ((sT[c + kBX] + sT[c - kBX] - 2.0 * T) * g.inv_dy2);
      ddT += g.kappa * ((Tn + Tm - 2.0 * T) * g.inv_dz2);
      const double x = ddT + sSp[tz * kTX * kTY] + g.S;
      double q = x * inv;                       // x / prescaler
      q = fma(fma(-prescaler, q, x), inv, q);
      double Tnew = T + q * g.inner_dt;
      if (Tnew < 0.0) {
        Tnew = 0.0;
        atomicOr(g.status, 2u);
      }
      out[tz * sz] = Tnew;
      if (g.clear_source) src[tz * sz] = 0.0;
    }
    Tm = T;
    T = Tn;
  }
}

}  // namespace ephb
