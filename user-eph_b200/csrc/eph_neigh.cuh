// Device-side construction of the full neighbour list (cut-off r_c + skin) from the positions, as an alternative to
// uploading LAMMPS' list (2.2 GB at 4 M atoms): cell binning with half-cut-off cells, counting pass, exclusive scan,
// fill pass.  Produces the same pair set LAMMPS' `REQ_FULL` list would (rows of local atoms over locals and ghosts).
#pragma once

#include "eph_device.cuh"

namespace ephb {

// Capacity of every warp tile of the inner list (eph_sweeps.cuh): 32 entries per iteration, as many iterations as
// the longest LAMMPS row of the tile needs (the inner list is a subset of LAMMPS' list), rounded up to a multiple of four.
// caps[ntiles] = 0 closes the scan.
__global__ void tile_caps_kernel(int nlocal, const long long *__restrict__ offsets, int tile_atoms, int lanes,
                                 int ntiles, long long *__restrict__ caps) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > ntiles) return;
  long long longest = 0;
  if (w < ntiles)
    for (int a = 0; a < tile_atoms; ++a) {
      const int i = w * tile_atoms + a;
      if (i < nlocal) longest = max(longest, offsets[i + 1] - offsets[i]);
    }
  // a multiple of four iterations: the packed sweeps pad every list to whole blocks of up to four (eph_packed.cuh)
  caps[w] = 32 * (((longest + lanes - 1) / lanes + 3) / 4 * 4);
}

// Boundary-first ordering of the density pass (multi-rank overlap): tiles that hold an atom some other rank needs as a
// ghost are flagged, then split into two ordered work lists with an exclusive scan of the flags.
__global__ void tile_mark_kernel(int n, const int *__restrict__ index, int tile_atoms, int nlocal, int *__restrict__ flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int a = index[t];
  if (a >= 0 && a < nlocal) flag[a / tile_atoms] = 1;
}
__global__ void tile_split_kernel(int ntiles, const int *__restrict__ flag, const int *__restrict__ scan,
                                  int *__restrict__ work /* boundary tiles, then interior tiles */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  if (flag[t]) work[scan[t]] = t;
  else work[scan[ntiles] + t - scan[t]] = t;
}

struct CellGrid {
  double lo[3];
  double inv[3];   // cells per length
  int nb[3];
};

__global__ void bbox_kernel(int n, const double *__restrict__ x, double *__restrict__ out /* lo[3], hi[3] as ordered bits */) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x)
    for (int d = 0; d < 3; ++d) {
      const double v = x[3 * (size_t)a + d];
      lo[d] = fmin(lo[d], v);
      hi[d] = fmax(hi[d], v);
    }
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    // atomic min/max on doubles through a CAS loop (set-up time only)
    for (int d = 0; d < 3; ++d) {
      unsigned long long *pl = reinterpret_cast<unsigned long long *>(out + d), *ph = reinterpret_cast<unsigned long long *>(out + 3 + d);
      unsigned long long old = *pl;
      while (__longlong_as_double((long long)old) > lo[d]) {
        const unsigned long long prev = atomicCAS(pl, old, (unsigned long long)__double_as_longlong(lo[d]));
        if (prev == old) break;
        old = prev;
      }
      old = *ph;
      while (__longlong_as_double((long long)old) < hi[d]) {
        const unsigned long long prev = atomicCAS(ph, old, (unsigned long long)__double_as_longlong(hi[d]));
        if (prev == old) break;
        old = prev;
      }
    }
  }
}

__device__ __forceinline__ int cell_coord(double v, double lo, double inv, int nb) {
  return min(nb - 1, max(0, static_cast<int>((v - lo) * inv)));
}

__global__ void cell_id_kernel(int n, const double *__restrict__ x, CellGrid g, int *__restrict__ cell, int *__restrict__ atom) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int cx = cell_coord(x[3 * (size_t)a], g.lo[0], g.inv[0], g.nb[0]);
  const int cy = cell_coord(x[3 * (size_t)a + 1], g.lo[1], g.inv[1], g.nb[1]);
  const int cz = cell_coord(x[3 * (size_t)a + 2], g.lo[2], g.inv[2], g.nb[2]);
  cell[a] = (cz * g.nb[1] + cy) * g.nb[0] + cx;
  atom[a] = a;
}

// positions in cell order (x, y, z, atom id) and the start of every cell's run
__global__ void cell_ranges_kernel(int n, const int *__restrict__ sorted_cell, const int *__restrict__ sorted_atom,
                                   const double *__restrict__ x, double4 *__restrict__ xs, int *__restrict__ cell_start,
                                   int *__restrict__ cell_end) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int c = sorted_cell[s], a = sorted_atom[s];
  xs[s] = make_double4(x[3 * (size_t)a], x[3 * (size_t)a + 1], x[3 * (size_t)a + 2], bits_to_double((unsigned)a));
  if (s == 0 || sorted_cell[s - 1] != c) cell_start[c] = s;
  if (s == n - 1 || sorted_cell[s + 1] != c) cell_end[c] = s + 1;
}

// FILL = false: count neighbours of local atom i; FILL = true: write them at offsets[i]
template <bool FILL>
__global__ void __launch_bounds__(128) neighbor_pass_kernel(int nlocal, const double *__restrict__ x, CellGrid g, double cut_sq,
                                                            const double4 *__restrict__ xs, const int *__restrict__ cell_start,
                                                            const int *__restrict__ cell_end, long long *__restrict__ counts,
                                                            const long long *__restrict__ offsets, int *__restrict__ neigh) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const double xi = x[3 * (size_t)i], yi = x[3 * (size_t)i + 1], zi = x[3 * (size_t)i + 2];
  const int cx = cell_coord(xi, g.lo[0], g.inv[0], g.nb[0]);
  const int cy = cell_coord(yi, g.lo[1], g.inv[1], g.nb[1]);
  const int cz = cell_coord(zi, g.lo[2], g.inv[2], g.nb[2]);
  long long n = 0;
  int *out = FILL ? neigh + offsets[i] : nullptr;
  for (int dz = -2; dz <= 2; ++dz) {
    const int z = cz + dz;
    if (z < 0 || z >= g.nb[2]) continue;
    for (int dy = -2; dy <= 2; ++dy) {
      const int y = cy + dy;
      if (y < 0 || y >= g.nb[1]) continue;
      // the five x cells of this row are consecutive cells: one contiguous run of the sorted array
      const int x0 = max(cx - 2, 0), x1 = min(cx + 2, g.nb[0] - 1);
      const int row = (z * g.nb[1] + y) * g.nb[0];
      int s0 = -1, s1 = -1;
      for (int c = row + x0; c <= row + x1; ++c) {
        const int b = cell_start[c], e = cell_end[c];
        if (e > b) {
          if (s0 < 0) s0 = b;
          s1 = e;
        }
      }
      for (int s = s0; s < s1; ++s) {
        const double4 p = xs[s];
        const int j = (int)double_to_bits(p.w);
        const double ddx = p.x - xi, ddy = p.y - yi, ddz = p.z - zi;
        if (j != i && ddx * ddx + ddy * ddy + ddz * ddz < cut_sq) {
          if (FILL) out[n] = j;
          ++n;
        }
      }
    }
  }
  if (!FILL) counts[i] = n;
}

// The same two passes, cooperatively: one CTA per row of cells (fixed y, z), its warps take chunks of 32 consecutive
// atoms of the row in cell order.  The atoms of a chunk share one candidate set -- the cells from two left of the
// chunk's first atom to two right of its last one, in the 25 rows around -- which the warp stages through shared memory
// 32 positions at a time (one coalesced 1 KB load) and every lane then reads as broadcasts.  The per-atom kernel above
// gathers ~1000 sectors per atom and pass (25 x 10 cell-range look-ups + ~490 candidates); here the memory side is ~75
// wavefronts per atom and the pair tests themselves (2.4 x as many, the price of the shared candidate set) set the pace,
// so they run in fp32 on coordinates relative to the chunk (error of r^2 below 3e-4 A^2) and only a pair inside a band
// of 1e-3 A^2 around a cut-off is decided again in fp64: the lists are exactly those of the fp64 test.
// Neighbours come out in the same order as from the per-atom kernel (rows by dz, dy; ascending cell order within a
// row).  Rows of ghost atoms are not built (the list has rows for local atoms only).
// INNER (fill pass only): the entries closer than r_c + inner_skin also go straight into the atom's slots of the inner
// list (warp-tiled layout of eph_sweeps.cuh, LANES lanes per atom), so a re-neighbouring needs no second walk of the
// new list; tile_pad_kernel completes the tiles afterwards.
constexpr int kNeighWarps = 4;
struct InnerOut {
  int *ineigh;                 // nullptr: no inner list
  int *icount;
  const long long *tile_off;
  const int *mask;             // atom->mask; atoms outside the fix group (mask & groupbit == 0) have no inner list
  int groupbit;
  double r_inner_sq;
  int lanes;
};
template <bool FILL>
__global__ void __launch_bounds__(32 * kNeighWarps) neighbor_tile_kernel(int nlocal, CellGrid g, double cut_sq, const double4 *__restrict__ xs,
                                                                         const int *__restrict__ cell_start, const int *__restrict__ cell_end,
                                                                         long long *__restrict__ counts, const long long *__restrict__ offsets,
                                                                         int *__restrict__ neigh, InnerOut io) {
  __shared__ double4 s_cand[kNeighWarps][32];
  __shared__ float4 s_rel[kNeighWarps][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int row = blockIdx.x;   // y + nb[1] * z
  const int ry = row % g.nb[1], rz = row / g.nb[1];
  const float band = 1e-3f;
  const float cut_lo = (float)cut_sq - band, cut_hi = (float)cut_sq + band;
  const float in_lo = (float)io.r_inner_sq - band, in_hi = (float)io.r_inner_sq + band;
  // the row's atoms: one contiguous run of the cell-sorted array (empty cells have start = end = 0)
  int r0 = -1, r1 = -1;
  for (int c = row * g.nb[0] + lane; c < (row + 1) * g.nb[0]; c += 32) {
    const int b = cell_start[c], e = cell_end[c];
    if (e > b) { r0 = (r0 < 0 || b < r0) ? b : r0; r1 = e > r1 ? e : r1; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int q0 = __shfl_xor_sync(0xFFFFFFFFu, r0, o), q1 = __shfl_xor_sync(0xFFFFFFFFu, r1, o);
    r0 = (r0 < 0) ? q0 : ((q0 >= 0 && q0 < r0) ? q0 : r0);
    r1 = q1 > r1 ? q1 : r1;
  }
  if (r0 < 0) return;
  for (int base = r0 + 32 * wib; base < r1; base += 32 * kNeighWarps) {
    const int sme = base + lane;
    const bool have = sme < r1;
    double4 me = make_double4(0, 0, 0, 0);
    int id = -1, cx = 0;
    if (have) {
      me = xs[sme];
      id = (int)double_to_bits(me.w);
      cx = cell_coord(me.x, g.lo[0], g.inv[0], g.nb[0]);
    }
    const bool mine = have && id < nlocal;   // ghosts have no row
    // x-cell range of the chunk (atoms are in cell order: first and last valid lane) and its origin for the fp32 tests
    const unsigned vmask = __ballot_sync(0xFFFFFFFFu, have);
    const int cxa = __shfl_sync(0xFFFFFFFFu, cx, 0), cxb = __shfl_sync(0xFFFFFFFFu, cx, 31 - __clz(vmask));
    const double ox = __shfl_sync(0xFFFFFFFFu, me.x, 0), oy = __shfl_sync(0xFFFFFFFFu, me.y, 0), oz = __shfl_sync(0xFFFFFFFFu, me.z, 0);
    const float mx = (float)(me.x - ox), my = (float)(me.y - oy), mz = (float)(me.z - oz);
    const int x0 = max(cxa - 2, 0), x1 = min(cxb + 2, g.nb[0] - 1);
    long long n = 0;
    int *out = (FILL && mine) ? neigh + offsets[id] : nullptr;
    const bool inner = FILL && mine && io.ineigh != nullptr && (io.mask[id] & io.groupbit) != 0;
    int ni = 0;
    long long it0 = 0;
    if (inner) {
      const int tile_atoms = 32 / io.lanes;
      it0 = io.tile_off[id / tile_atoms] + (long long)(id % tile_atoms) * io.lanes;
    }
    for (int dz = -2; dz <= 2; ++dz) {
      const int z = rz + dz;
      if (z < 0 || z >= g.nb[2]) continue;
      for (int dy = -2; dy <= 2; ++dy) {
        const int y = ry + dy;
        if (y < 0 || y >= g.nb[1]) continue;
        const int crow = (z * g.nb[1] + y) * g.nb[0];
        // the run of the cells [x0, x1] of that row
        int s0 = -1, s1 = -1;
        for (int c = crow + x0 + lane; c <= crow + x1; c += 32) {
          const int b = cell_start[c], e = cell_end[c];
          if (e > b) { s0 = (s0 < 0 || b < s0) ? b : s0; s1 = e > s1 ? e : s1; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const int q0 = __shfl_xor_sync(0xFFFFFFFFu, s0, o), q1 = __shfl_xor_sync(0xFFFFFFFFu, s1, o);
          s0 = (s0 < 0) ? q0 : ((q0 >= 0 && q0 < s0) ? q0 : s0);
          s1 = q1 > s1 ? q1 : s1;
        }
        if (s0 < 0) continue;
        // the whole run is tested against every atom of the chunk: outside an atom's own window of cells [cx - 2, cx + 2]
        // the distance test fails anyway (cells are at least cut-off / 2 wide), so the rows equal the per-atom kernel's
        for (int cb = s0; cb < s1; cb += 32) {
          __syncwarp();
          if (cb + lane < s1) {
            const double4 p = xs[cb + lane];
            s_cand[wib][lane] = p;
            s_rel[wib][lane] = make_float4((float)(p.x - ox), (float)(p.y - oy), (float)(p.z - oz), __int_as_float((int)double_to_bits(p.w)));
          }
          __syncwarp();
          const int m = min(32, s1 - cb);
          if (mine) {
            for (int c = 0; c < m; ++c) {
              const float4 q = s_rel[wib][c];
              const float fx = q.x - mx, fy = q.y - my, fz = q.z - mz;
              const float d2 = fx * fx + fy * fy + fz * fz;
              if (d2 >= cut_hi) continue;
              const int j = __float_as_int(q.w);
              if (j == id) continue;
              bool in_cut = d2 < cut_lo, in_inner = d2 < in_lo;
              if (!in_cut || (!in_inner && d2 < in_hi)) {   // inside a band: the fp64 test decides
                const double4 p = s_cand[wib][c];
                const double ddx = p.x - me.x, ddy = p.y - me.y, ddz = p.z - me.z;
                const double r2 = ddx * ddx + ddy * ddy + ddz * ddz;
                in_cut = r2 < cut_sq;
                in_inner = r2 < io.r_inner_sq;
              }
              if (!in_cut) continue;
              if (FILL) out[n] = j;
              ++n;
              if (inner && in_inner) {
                io.ineigh[it0 + (long long)(ni / io.lanes) * 32 + (ni % io.lanes)] = j;
                ++ni;
              }
            }
          }
        }
      }
    }
    if (!FILL && mine) counts[id] = n;
    if (FILL && mine && io.ineigh != nullptr) io.icount[id] = ni;
  }
}

// Completes freshly written tiles of the inner list: every atom's slots from its list length up to the longest list of
// the tile, rounded up to kPadIters iterations, get the atom's own index and a zero pair weight (see eph_packed.cuh).
// One warp per tile.
__global__ void __launch_bounds__(256) tile_pad_kernel(int nlocal, int lanes, const long long *__restrict__ tile_off,
                                                       const int *__restrict__ icount, int *__restrict__ ineigh,
                                                       double *__restrict__ gpair, double *__restrict__ gpair_i) {
  const int lane = threadIdx.x & 31;
  const int tile_atoms = 32 / lanes;
  const int ntiles = (nlocal + tile_atoms - 1) / tile_atoms;
  for (int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < ntiles; tile += gridDim.x * (blockDim.x >> 5)) {
    const int i_mine = tile * tile_atoms + lane;
    const int cnt_mine = (lane < tile_atoms && i_mine < nlocal) ? icount[i_mine] : 0;
    int tmax = cnt_mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmax = max(tmax, __shfl_xor_sync(0xFFFFFFFFu, tmax, o));
    const int padded = (tmax + 4 * lanes - 1) / (4 * lanes) * (4 * lanes);   // kPadIters = 4 iterations (eph_sweeps.cuh)
    const long long t0 = tile_off[tile];
    for (int t = 0; t < tile_atoms; ++t) {
      const int i = tile * tile_atoms + t;
      const int cnt = __shfl_sync(0xFFFFFFFFu, cnt_mine, t);
      for (int c = cnt + lane; c < padded; c += 32) {
        const long long dst = t0 + (long long)(c / lanes) * 32 + t * lanes + (c % lanes);
        ineigh[dst] = i < nlocal ? i : 0;
        gpair[dst] = 0.0;
        if (gpair_i != nullptr) gpair_i[dst] = 0.0;
      }
    }
  }
}

}  // namespace ephb
